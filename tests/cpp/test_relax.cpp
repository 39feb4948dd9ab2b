// test/test_relax.cpp:76-166 against the shim: RELAX with k = 1 gives the same likelihood whichever branches carry which copy,
// the same as the site model YNGP_M2, and RELAX with k = 2 on one group equals two YNGP_M2 copies with the induced omegas.
#include <cmath>
#include <cstdio>
#include <iostream>
#include <memory>

#include "../../bpp_phyl_b200/host/bppgpu_shim.hpp"

using namespace bppshim;
using namespace std;

int main() {
  int fails = 0;
  try {
    unique_ptr<Tree> tree(TreeTemplateTools::parenthesisToTree("(((A:0.01, B:0.01):0.02,C:0.03):0.01,D:0.04);"));
    const CodonAlphabet* alphabet = &AlphabetTools::CODON_ALPHABET();
    VectorSiteContainer sites(alphabet);
    sites.addSequence(BasicSequence("A", "AAATGGCTGTGCACGTCT", alphabet));
    sites.addSequence(BasicSequence("B", "AACTGGATCTGCATGTCT", alphabet));
    sites.addSequence(BasicSequence("C", "ATCTGGACGTGCACGTGT", alphabet));
    sites.addSequence(BasicSequence("D", "CAACGGGAGTGCGCCTAT", alphabet));
    ConstantRateDistribution rdist;

    // partition A: model1 on node 0, model2 on nodes 1..5 (RELAX, k = 1, parameters of model2 aliased to model1's)
    MixedSubstitutionModelSet relax1(alphabet);
    relax1.addModel(new RELAX(alphabet, 2.0, 0.1, 1.0, 2.0, 1.0, 0.5, 0.8), {0});
    relax1.addModel(new RELAX(alphabet, 2.0, 0.1, 1.0, 2.0, 1.0, 0.5, 0.8), {1, 2, 3, 4, 5});
    RNonHomogeneousMixedTreeLikelihood tl1(*tree, sites, &relax1, &rdist, true, false);
    tl1.initialize();
    const double l1 = -tl1.getValue();
    printf("RELAX_PARTITION_A %.15f\n", l1);

    // partition B
    MixedSubstitutionModelSet relax2(alphabet);
    relax2.addModel(new RELAX(alphabet, 2.0, 0.1, 1.0, 2.0, 1.0, 0.5, 0.8), {1, 2, 3, 4, 5});
    relax2.addModel(new RELAX(alphabet, 2.0, 0.1, 1.0, 2.0, 1.0, 0.5, 0.8), {0});
    RNonHomogeneousMixedTreeLikelihood tl2(*tree, sites, &relax2, &rdist, true, false);
    tl2.initialize();
    const double l2 = -tl2.getValue();
    printf("RELAX_PARTITION_B %.15f\n", l2);
    if (fabs(l1 - l2) > 0.001) { cout << "Error! different likelihood is computed in RELAX for different partitions when k=1" << endl; fails++; }

    // the site model YNGP_M2
    YNGP_M2 m2(alphabet, 2.0, 0.1, 2.0, 0.5, 0.8);
    RHomogeneousMixedTreeLikelihood tlm2(*tree, sites, &m2, &rdist, true, false);
    tlm2.initialize();
    const double lm2 = -tlm2.getValue();
    printf("M2 %.15f\n", lm2);
    for (size_t k = 0; k < 3; ++k) printf("M2_RATE_%zu %.15g\n", k, m2.getNModel(k)->getRate());
    for (size_t k = 0; k < 3; ++k) printf("M2_PROB_%zu %.15g\n", k, m2.getNProbability(k));
    if (fabs(l1 - lm2) > 0.001) { cout << "Error! RELAX when k=1 yields different likelihood than M2 model" << endl; fails++; }

    // k = 2 on model2 of partition B = two YNGP_M2 copies with the induced omegas
    tl2.setParameterValue("RELAX.k_2", 2);
    tl2.computeTreeLikelihood();
    const double l3 = -tl2.getValue();
    printf("RELAX_K2 %.15f\n", l3);
    MixedSubstitutionModelSet doubleM2(alphabet);
    doubleM2.addModel(new YNGP_M2(alphabet, 2.0, 0.1, 2.0, 0.5, 0.8), {1, 2, 3, 4, 5});
    doubleM2.addModel(new YNGP_M2(alphabet, 2.0, 0.01, 4.0, 0.5, 0.8), {0});
    RNonHomogeneousMixedTreeLikelihood tld(*tree, sites, &doubleM2, &rdist, true, false);
    tld.initialize();
    const double ld = -tld.getValue();
    printf("DOUBLE_M2 %.15f\n", ld);
    if (fabs(l3 - ld) > 0.001) { cout << "Error! RELAX yields different likelihood from two copies of YNGP_M2" << endl; fails++; }
    if (fabs(l3 - l1) < 1e-6) { cout << "Error! k = 2 did not change the likelihood" << endl; fails++; }
  } catch (std::exception& e) {
    cerr << e.what() << endl;
    return 1;
  }
  if (fails) cerr << fails << " check(s) failed" << endl;
  return fails ? 1 : 0;
}
