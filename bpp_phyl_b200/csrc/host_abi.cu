// Two small C-ABI utilities that are not part of the evaluation path:
//   bppgpu_host_model         the C++ shim's model classes (updateMatrices on the HOST, Model/AbstractSubstitutionModel.cpp:
//                             175-421 and the per-model generators) reachable from a plain-C / ctypes harness, so that benches
//                             and tests build LG08 / YN98 / GY94 / GTR / Chromosome eigensystems with the product's own host code;
//   bppgpu_measure_fp64_peak  the FP64 ceilings of the device measured in place (DFMA on the CUDA-core FP64 pipe, DMMA
//                             m8n8k4 on the tensor path), so a roofline fraction is quoted against a peak taken in the same
//                             run under the same clocks.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/bppgpu.h"
#include "../host/bppgpu_shim.hpp"
#include "common.cuh"

using namespace bppgpu;

namespace {

__global__ void peak_dfma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void peak_dmma_kernel(double* out, int iters) {
  double c[8][2] = {};
  const double a = threadIdx.x * 1e-3, b = 1e-3;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[j][0]), "+d"(c[j][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

int bppgpu_measure_fp64_peak(int device, double* dfma_tflops, double* dmma_tflops) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) BPP_FAIL(BPPGPU_E_CUDA, "no usable CUDA device; libbppgpu has no CPU fallback");
  if (device < 0 || device >= n) BPP_FAIL(BPPGPU_E_INVALID, "device %d out of range", device);
  BPP_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  BPP_CUDA(cudaGetDeviceProperties(&prop, device));
  const int grid = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
  double* out = nullptr;
  BPP_CUDA(cudaMalloc(&out, (size_t)grid * threads * 8));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto best_of = [&](auto launch) {
    launch();
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
      cudaEventRecord(e0);
      launch();
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
    }
    return (double)best;
  };
  const double ms_f = best_of([&] { peak_dfma_kernel<<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
  const double ms_m = best_of([&] { peak_dmma_kernel<<<grid, threads>>>(out, iters); });
  cudaError_t err = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  BPP_CUDA(err);
  if (dfma_tflops) *dfma_tflops = (double)grid * threads * iters * 8.0 * 2.0 / (ms_f * 1e-3) / 1e12;
  if (dmma_tflops) *dmma_tflops = (double)grid * (threads / 32) * iters * 8.0 * (2.0 * 8 * 8 * 4) / (ms_m * 1e-3) / 1e12;
  return BPPGPU_OK;
}

int bppgpu_host_model(const char* name, const double* params, int32_t n_params, int32_t* n_states, uint32_t* flags, double* rate,
                      double* Q, double* V, double* Vinv, double* eigen_re, double* eigen_im, double* freq) {
  if (!name || !n_states) BPP_FAIL(BPPGPU_E_INVALID, "null name / n_states");
  if (n_params > 0 && !params) BPP_FAIL(BPPGPU_E_INVALID, "null params");
  using namespace bppshim;
  auto P = [&](int i, double def) { return i < n_params ? params[i] : def; };
  try {
    std::unique_ptr<SubstitutionModel> m;
    std::unique_ptr<ChromosomeAlphabet> chr;
    const std::string nm(name);
    const DNA* dna = &AlphabetTools::DNA_ALPHABET();
    if (nm == "GTR") m.reset(new GTR(dna, P(0, 1), P(1, 1), P(2, 1), P(3, 1), P(4, 1), P(5, .25), P(6, .25), P(7, .25), P(8, .25)));
    else if (nm == "HKY85") m.reset(new HKY85(dna, P(0, 1), P(1, .25), P(2, .25), P(3, .25), P(4, .25)));
    else if (nm == "T92") m.reset(new T92(dna, P(0, 1), P(1, .5)));
    else if (nm == "K80") m.reset(new K80(dna, P(0, 1)));
    else if (nm == "JC69") m.reset(new JCnuc(dna));
    else if (nm == "LG08") m.reset(new LG08(&AlphabetTools::PROTEIN_ALPHABET()));
    else if (nm == "YN98") m.reset(new YN98(&AlphabetTools::CODON_ALPHABET(), P(0, 1), P(1, 1)));
    else if (nm == "GY94") m.reset(new GY94(&AlphabetTools::CODON_ALPHABET(), P(0, 1), P(1, 10000)));
    else if (nm == "Chromosome") {
      // params: min, max, gain, loss, dupl, demi [, gainR, lossR, duplR, baseNum, baseNumR, maxChrRange]
      if (n_params < 6) BPP_FAIL(BPPGPU_E_INVALID, "Chromosome needs min, max, gain, loss, dupl, demi");
      chr.reset(new ChromosomeAlphabet((unsigned)params[0], (unsigned)params[1]));
      const double ig = ChromosomeSubstitutionModel::IgnoreParam;
      m.reset(new ChromosomeSubstitutionModel(chr.get(), params[2], params[3], params[4], params[5], P(6, ig), P(7, ig), P(8, ig),
                                              (int)P(9, ig), P(10, ig), (unsigned)P(11, 0)));
    } else {
      BPP_FAIL(BPPGPU_E_INVALID, "unknown model '%s'", name);
    }
    const int S = (int)m->getNumberOfStates();
    *n_states = S;
    bppgpu_model_desc d{};
    m->fillModelDesc(d);
    if (flags) *flags = d.flags;
    if (rate) *rate = d.rate;
    const size_t SS = (size_t)S * S;
    if (Q) memcpy(Q, d.generator, SS * 8);
    if (V) memcpy(V, d.right_eigen, SS * 8);
    if (Vinv) memcpy(Vinv, d.left_eigen, SS * 8);
    if (eigen_re) memcpy(eigen_re, d.eigen_re, S * 8);
    if (eigen_im) memcpy(eigen_im, d.eigen_im, S * 8);
    if (freq) memcpy(freq, m->getFrequencies().data(), S * 8);
  } catch (std::exception& ex) {
    BPP_FAIL(BPPGPU_E_INVALID, "%s", ex.what());
  }
  return BPPGPU_OK;
}

// ---- a4: the R classes' recursive per-subtree compression (bit-exact integer work, host side) -------------------------
// DRASRTreeLikelihoodData::initLikelihoodsWithPatterns (Likelihood/DRASRTreeLikelihoodData.cpp:218-332): at every node the
// incoming container (the father's unique columns) is cut down to the node's own leaves (PatternTools::getSequenceSubset,
// PatternTools.cpp:59-70: the leaves in tree order), compressed again (SitePatterns, SitePatterns.cpp:52-106) and handed to the
// sons; the son's `indices_` is patternLinks_[father][son] (:323-325, DRASRTreeLikelihoodData.h:145).
namespace {
struct SubtreeCtx {
  int elem;                       // bytes per sequence element
  const int32_t *child_off, *children, *leaf_seq;
  int64_t* n_patterns;
  std::vector<std::vector<int64_t>> links;   // per node: father pattern -> this node's pattern
};
// unique sorted columns of `cols` (each `w` bytes); idx[i] = pattern of column i; weights optional
static void compress_columns(const std::vector<std::string>& cols, std::vector<std::string>& uniq, std::vector<int64_t>& idx,
                             std::vector<uint32_t>* weights) {
  const size_t n = cols.size();
  std::vector<int64_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
    const int c = cols[(size_t)a].compare(cols[(size_t)b]);
    return c != 0 ? c < 0 : a < b;
  });
  uniq.clear();
  idx.assign(n, 0);
  if (weights) weights->clear();
  for (size_t k = 0; k < n; ++k) {
    const std::string& c = cols[(size_t)order[k]];
    if (k == 0 || c != uniq.back()) {
      uniq.push_back(c);
      if (weights) weights->push_back(0);
    }
    if (weights) weights->back()++;
    idx[(size_t)order[k]] = (int64_t)uniq.size() - 1;
  }
}
static void leaves_of(const SubtreeCtx& cx, int node, std::vector<int>& out) {
  if (cx.child_off[node + 1] == cx.child_off[node]) { out.push_back(node); return; }
  for (int k = cx.child_off[node]; k < cx.child_off[node + 1]; ++k) leaves_of(cx, cx.children[k], out);
}
// `order`: the leaf nodes whose elements make up a column of `cols`, in column order
static void subtree_rec(SubtreeCtx& cx, int node, const std::vector<int>& order, const std::vector<std::string>& cols,
                        std::vector<int64_t>& idx_out, std::vector<uint32_t>* weights) {
  std::vector<int> lv;
  leaves_of(cx, node, lv);
  std::vector<int> pos(lv.size());
  for (size_t k = 0; k < lv.size(); ++k) pos[k] = (int)(std::find(order.begin(), order.end(), lv[k]) - order.begin());
  std::vector<std::string> sub(cols.size());
  for (size_t i = 0; i < cols.size(); ++i) {
    sub[i].reserve(lv.size() * cx.elem);
    for (int p : pos) sub[i].append(cols[i], (size_t)p * cx.elem, (size_t)cx.elem);
  }
  std::vector<std::string> uniq;
  compress_columns(sub, uniq, idx_out, weights);
  cx.n_patterns[node] = (int64_t)uniq.size();
  for (int k = cx.child_off[node]; k < cx.child_off[node + 1]; ++k) {
    const int son = cx.children[k];
    subtree_rec(cx, son, lv, uniq, cx.links[(size_t)son], nullptr);
  }
}
}  // namespace

int bppgpu_subtree_patterns(const uint8_t* columns, int64_t n_sites, int32_t n_seqs, int32_t elem_bytes, int32_t n_nodes,
                            const int32_t* child_offsets, const int32_t* children, int32_t root, const int32_t* leaf_seq,
                            int64_t* n_patterns, int64_t* link_offsets, int64_t* links, int64_t* root_links,
                            uint32_t* root_weights) {
  if (n_sites < 0 || n_seqs <= 0 || elem_bytes <= 0 || n_nodes <= 0 || !child_offsets || !children || !leaf_seq || !n_patterns ||
      !link_offsets || (n_sites > 0 && !columns) || root < 0 || root >= n_nodes)
    BPP_FAIL(BPPGPU_E_INVALID, "bad argument to bppgpu_subtree_patterns");
  for (int n = 0; n < n_nodes; ++n) {
    const bool leaf = child_offsets[n + 1] == child_offsets[n];
    if (leaf && (leaf_seq[n] < 0 || leaf_seq[n] >= n_seqs)) BPP_FAIL(BPPGPU_E_INVALID, "leaf node %d has no sequence", n);
  }
  SubtreeCtx cx{elem_bytes, child_offsets, children, leaf_seq, n_patterns, {}};
  cx.links.resize((size_t)n_nodes);
  // the incoming container of the root: every sequence, in container order; pseudo leaf ids = -(seq + 1) are resolved through
  // `order0`, which lists the leaf NODE of every sequence position (sequences without a leaf never match)
  std::vector<int> order0((size_t)n_seqs, -1);
  for (int n = 0; n < n_nodes; ++n)
    if (child_offsets[n + 1] == child_offsets[n]) order0[(size_t)leaf_seq[n]] = n;
  const size_t w = (size_t)n_seqs * elem_bytes;
  std::vector<std::string> cols((size_t)n_sites);
  for (int64_t i = 0; i < n_sites; ++i) cols[(size_t)i].assign((const char*)columns + (size_t)i * w, w);
  std::vector<int64_t> ridx;
  std::vector<uint32_t> rw;
  subtree_rec(cx, root, order0, cols, ridx, &rw);
  if (root_links) std::copy(ridx.begin(), ridx.end(), root_links);
  if (root_weights) std::copy(rw.begin(), rw.end(), root_weights);
  int64_t off = 0;
  for (int n = 0; n < n_nodes; ++n) {
    link_offsets[n] = off;
    off += (int64_t)cx.links[(size_t)n].size();
  }
  link_offsets[n_nodes] = off;
  if (links)
    for (int n = 0; n < n_nodes; ++n) std::copy(cx.links[(size_t)n].begin(), cx.links[(size_t)n].end(), links + link_offsets[n]);
  return BPPGPU_OK;
}
