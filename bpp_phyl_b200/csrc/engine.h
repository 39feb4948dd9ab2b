// Host-side engine state behind the opaque bppgpu_engine handle.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/bppgpu.h"
#include "pt_kernels.cuh"
#include "walk_kernels.cuh"

namespace bppgpu {

enum PathKind { PATH_NONE = 0, PATH_WALK4 = 1, PATH_WALKS = 2, PATH_GENERIC = 3, PATH_DMMA = 4, PATH_POINTS = 5 };

struct DevModel {
  double *V = nullptr, *Vinv = nullptr, *re = nullptr, *im = nullptr, *Q = nullptr, *Q2 = nullptr;
  int* role = nullptr;
  double rate = 1.0, eps = 1e-4;
  unsigned flags = 0;
  int has_complex = 0;
  bool set = false;
  double q_max_abs_diag = 0.0, q_l1 = 0.0;
};

struct Program {
  std::vector<Op> ops;
  std::vector<Child> childs;
  int nslots = 0;
  Op* d_ops = nullptr;
  Child* d_childs = nullptr;
  bool built = false;
};

}  // namespace bppgpu

struct bppgpu_engine {
  int dev = 0;
  cudaStream_t stream = nullptr;
  int S = 0, C = 0, nn = 0, root = 0, npoints = 1, nmodels = 1, ncodes = 0, code_bytes = 1;
  long long N = 0;
  unsigned flags = 0;
  // topology (host)
  std::vector<int> child_off, children, parent, leaf_slot, leaf_nodes, internal_idx;
  int nl = 0, ni = 0;
  // device inputs
  void* d_codes = nullptr;
  double* d_code_table = nullptr;
  double* d_weights = nullptr;
  double *d_rates = nullptr, *d_probs = nullptr;
  double* d_rootfreq = nullptr;       // [npoints][S]
  double* d_rootfreq_used = nullptr;  // [npoints][S]
  double* d_brlen = nullptr;          // [npoints][nn]
  int* d_branch_model = nullptr;      // [npoints][nn]
  int* d_leaf_nodes = nullptr;
  std::vector<bppgpu::DevModel> models;
  bppgpu::ModelDev* d_models = nullptr;
  bool models_dirty = true;
  std::vector<double> h_rates, h_probs;
  std::vector<double> h_brlen;  // [npoints][nn]
  std::vector<int> h_branch_model;
  // tables
  double *d_P = nullptr, *d_dP = nullptr, *d_d2P = nullptr;
  double* d_tiptab = nullptr;
  // CLV storage
  double* d_keep = nullptr;  // [ni][N][C][S]
  int* d_keep_exp = nullptr;
  double* d_gstack = nullptr;
  int* d_gstack_exp = nullptr;
  double* d_upper = nullptr;  // derivative pass scratch
  int* d_upper_exp = nullptr;
  int upper_slots = 0;
  // outputs
  double* d_SR = nullptr;
  int* d_rexp = nullptr;
  double* d_site_lnl = nullptr;  // [npoints][N]
  double* d_partials = nullptr;
  int n_partials = 0;
  double* d_out = nullptr;  // [npoints][1+2nn]
  double* d_deriv_partials = nullptr;
  // schedule
  bppgpu::Program prog;
  bppgpu::Program uprog;  // prefix/derivative program
  int path = bppgpu::PATH_NONE;
  // state flags
  bool have_weights = false, have_rates = false, have_brlen = false, have_rootfreq = false;
  std::vector<char> have_tip;
  int cached_point = -1;
  unsigned cached_want = 0;
  // stats
  bppgpu_stats stats{};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
  size_t bytes_resident = 0;
};
