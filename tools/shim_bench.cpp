// Timing of the C++ host path (the stated host product): setParameterValue -> getValue / derivatives through the shim classes, on a
// synthetic DNA data set (GTR + Gamma4, random tree, columns evolved crudely from an ancestral character so that patterns repeat).
// Prints one JSON line.   tools/shim_bench [n_taxa] [n_sites] [iterations]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "../bpp_phyl_b200/host/bppgpu_shim.hpp"

using namespace bppshim;
using namespace std;

static string random_newick(int n, mt19937_64& rng) {
  vector<string> pool;
  exponential_distribution<double> bl(1.0 / 0.05);
  for (int i = 0; i < n; ++i) pool.push_back("t" + to_string(i) + ":" + to_string(bl(rng) + 1e-4));
  while (pool.size() > 3) {
    uniform_int_distribution<size_t> pick(0, pool.size() - 1);
    size_t a = pick(rng), b = pick(rng);
    if (a == b) continue;
    if (a > b) swap(a, b);
    const string j = "(" + pool[a] + "," + pool[b] + "):" + to_string(bl(rng) + 1e-4);
    pool.erase(pool.begin() + b);
    pool[a] = j;
  }
  return "(" + pool[0] + "," + pool[1] + "," + pool[2] + ");";
}

int main(int argc, char** argv) {
  const int ntaxa = argc > 1 ? atoi(argv[1]) : 256, nsites = argc > 2 ? atoi(argv[2]) : 100000, iters = argc > 3 ? atoi(argv[3]) : 20;
  try {
    mt19937_64 rng(20260110);
    unique_ptr<Tree> tree(TreeTemplateTools::parenthesisToTree(random_newick(ntaxa, rng)));
    const DNA* dna = &AlphabetTools::DNA_ALPHABET();
    VectorSiteContainer sites(dna);
    const char* acgt = "ACGT";
    vector<int> anc(nsites);
    for (int& a : anc) a = (int)(rng() & 3);
    for (int t = 0; t < ntaxa; ++t) {
      string s(nsites, 'A');
      for (int i = 0; i < nsites; ++i) s[i] = acgt[(rng() % 100) < 12 ? (int)(rng() & 3) : anc[i]];
      sites.addSequence(BasicSequence("t" + to_string(t), s, dna));
    }
    GTR model(dna, 1.2, 0.8, 0.6, 1.5, 0.9, .3, .2, .25, .25);
    GammaDiscreteRateDistribution rdist(4, 0.5);
    auto t0 = chrono::steady_clock::now();
    DRHomogeneousTreeLikelihood tl(*tree, sites, &model, &rdist, true, false);
    tl.initialize();
    const double setup_s = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
    const double v0 = tl.getValue();
    auto timeit = [&](auto f) {
      f(0);
      auto a = chrono::steady_clock::now();
      for (int i = 1; i <= iters; ++i) f(i);
      return 1e3 * chrono::duration<double>(chrono::steady_clock::now() - a).count() / iters;
    };
    double sink = 0;
    const double ms_value = timeit([&](int i) { tl.setParameterValue("BrLen3", 0.05 + 1e-4 * i); sink += tl.getValue(); });
    const double ms_model = timeit([&](int i) { tl.setParameterValue("GTR.a", 1.2 + 1e-3 * i); sink += tl.getValue(); });
    const double ms_deriv = timeit([&](int i) {
      tl.setParameterValue("BrLen5", 0.04 + 1e-4 * i);
      sink += tl.getFirstOrderDerivative("BrLen5") + tl.getSecondOrderDerivative("BrLen7");
    });
    const double ms_site = timeit([&](int i) { tl.setParameterValue("BrLen2", 0.03 + 1e-4 * i); sink += tl.getLogLikelihoodForASite(i % nsites); });
    printf("{\"what\": \"C++ shim host path, DRHomogeneousTreeLikelihood, GTR+G4\", \"taxa\": %d, \"sites\": %d, \"patterns\": %zu, "
           "\"setup_s\": %.3f, \"minus_lnl\": %.6f, \"ms_setBrLen_getValue\": %.3f, \"ms_setModelParam_getValue\": %.3f, "
           "\"ms_setBrLen_d1_d2_all_branches\": %.3f, \"ms_setBrLen_getSiteLogLikelihood\": %.3f, \"iterations\": %d, \"sink\": %.3f}\n",
           ntaxa, nsites, tl.getNumberOfDistinctSites(), setup_s, v0, ms_value, ms_model, ms_deriv, ms_site, iters, sink);
  } catch (std::exception& e) {
    fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
