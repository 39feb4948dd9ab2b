// DRHomogeneousMixedTreeLikelihood (Likelihood/DRHomogeneousMixedTreeLikelihood.cpp) through the shim on the device: the site model
// YNGP_M2 on test/test_relax.cpp's data -- value (must equal the R class's), first and second branch derivatives, a branch
// move and a model-parameter move.  tests/test_cpp_shim.py re-derives every number with the oracle.
#include <cmath>
#include <cstdio>
#include <iostream>
#include <memory>

#include "../../bpp_phyl_b200/host/bppgpu_shim.hpp"

using namespace bppshim;
using namespace std;

int main() {
  try {
    unique_ptr<Tree> tree(TreeTemplateTools::parenthesisToTree("(((A:0.01, B:0.01):0.02,C:0.03):0.01,D:0.04);"));
    const CodonAlphabet* alphabet = &AlphabetTools::CODON_ALPHABET();
    VectorSiteContainer sites(alphabet);
    sites.addSequence(BasicSequence("A", "AAATGGCTGTGCACGTCT", alphabet));
    sites.addSequence(BasicSequence("B", "AACTGGATCTGCATGTCT", alphabet));
    sites.addSequence(BasicSequence("C", "ATCTGGACGTGCACGTGT", alphabet));
    sites.addSequence(BasicSequence("D", "CAACGGGAGTGCGCCTAT", alphabet));
    ConstantRateDistribution rdist;
    YNGP_M2 m2(alphabet, 2.0, 0.1, 2.0, 0.5, 0.8);
    DRHomogeneousMixedTreeLikelihood dr(*tree, sites, &m2, &rdist, /*checkRooted=*/false, false);
    dr.initialize();
    RHomogeneousMixedTreeLikelihood r(*tree, sites, &m2, &rdist, true, false);
    r.initialize();
    printf("DRM_VALUE %.15f\n", dr.getValue());
    printf("RM_VALUE %.15f\n", r.getValue());
    const ParameterList bl = dr.getBranchLengthsParameters();
    printf("DRM_NBRANCH %zu\n", bl.size());
    for (size_t b = 0; b < bl.size(); ++b) {
      printf("DRM_D1_%zu %.15g\n", b, dr.getFirstOrderDerivative(bl[b].name));
      printf("DRM_D2_%zu %.15g\n", b, dr.getSecondOrderDerivative(bl[b].name));
    }
    for (size_t i = 0; i < dr.getNumberOfSites(); ++i) printf("DRM_SITE_%zu %.15g\n", i, dr.getLogLikelihoodForASite(i));
    dr.setParameterValue("BrLen1", 0.2);
    printf("DRM_VALUE_MOVED %.15f\n", dr.getValue());
    printf("DRM_D1_MOVED_1 %.15g\n", dr.getFirstOrderDerivative("BrLen1"));
    m2.setParameterValue("omega0", 0.3);
    dr.modelChanged();
    printf("DRM_VALUE_OMEGA %.15f\n", dr.getValue());
    bool threw = false;
    try { dr.getFirstOrderDerivative("YNGP_M2.kappa"); } catch (Exception&) { threw = true; }
    printf("DRM_MODEL_DERIV_THROWS %d\n", (int)threw);
  } catch (std::exception& e) {
    cerr << e.what() << endl;
    return 1;
  }
  return 0;
}
