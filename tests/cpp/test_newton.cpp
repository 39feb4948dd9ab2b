// Branch-length optimisation by the reference's PseudoNewtonOptimizer (Likelihood/PseudoNewtonOptimizer.cpp:100-193, as
// OptimizationTools::optimizeNumericalParameters sets it on the branch lengths, OptimizationTools.cpp:187-188) driven by the
// device's first and second derivatives.  Prints the trajectory; tests/test_cpp_shim.py runs the same algorithm on the oracle.
#include <cmath>
#include <cstdio>
#include <iostream>
#include <memory>

#include "../../bpp_phyl_b200/host/bppgpu_shim.hpp"

using namespace bppshim;
using namespace std;

template <class TL>
static void run(const char* tag, TL& tl, double tol) {
  tl.initialize();
  PseudoNewtonOptimizer opt(&tl);
  opt.setTolerance(tol);
  opt.init(tl.getBranchLengthsParameters());
  printf("%s_START %.15f\n", tag, opt.getFunctionValue());
  int steps = 0;
  double prev;
  do {
    prev = opt.getFunctionValue();
    const double v = opt.step();
    printf("%s_STEP_%d %.15f\n", tag, steps, v);
    ++steps;
  } while (fabs(opt.getFunctionValue() - prev) >= tol && steps < 200);
  printf("%s_NSTEPS %d\n", tag, steps);
  printf("%s_NEVAL %u\n", tag, opt.getNumberOfEvaluations());
  printf("%s_FINAL %.15f\n", tag, opt.getFunctionValue());
  for (const Parameter& p : opt.getParameters()) printf("%s_%s %.15g\n", tag, p.name.c_str(), p.value);
  double g = 0;
  for (const Parameter& p : opt.getParameters())
    if (p.value > 1e-6 * 1.0001) g = max(g, fabs(tl.getFirstOrderDerivative(p.name)));
  printf("%s_MAXGRAD_INTERIOR %.6e\n", tag, g);
}

int main() {
  try {
    {
      const DNA dna;
      unique_ptr<Tree> t(TreeTemplateTools::parenthesisToTree("((A:0.01, B:0.02):0.03,C:0.01,D:0.1);"));
      VectorSiteContainer s(&dna);
      s.addSequence(BasicSequence("A", "AAATGGCTGTGCACGTC", &dna));
      s.addSequence(BasicSequence("B", "GACTGGATCTGCACGTC", &dna));
      s.addSequence(BasicSequence("C", "CTCTGGATGTGCACGTG", &dna));
      s.addSequence(BasicSequence("D", "AAATGGCGGTGCGCCTA", &dna));
      T92 m(&dna, 3.);
      GammaDiscreteRateDistribution g(4, 1.0);
      DRHomogeneousTreeLikelihood tl(*t, s, &m, &g);
      run("PN_T92", tl, 1e-6);
    }
    {
      const ProteicAlphabet* prot = &AlphabetTools::PROTEIN_ALPHABET();
      unique_ptr<Tree> t(TreeTemplateTools::parenthesisToTree("((a:0.1,b:0.2):0.05,(c:0.3,d:0.02):0.07,e:0.15);"));
      VectorSiteContainer s(prot);
      s.addSequence(BasicSequence("a", "ARNDCQEGHILKMFPSTWYVAAX", prot));
      s.addSequence(BasicSequence("b", "ARNDCQEGHILKMFPSTWYVLK-", prot));
      s.addSequence(BasicSequence("c", "ARNECQDGHLIKMFPTSWYVAKB", prot));
      s.addSequence(BasicSequence("d", "GRNDCQEGHILRMYPSTWFVAAZ", prot));
      s.addSequence(BasicSequence("e", "ARNDCQEGHVLKMFPSTWYIVAA", prot));
      LG08 lg(prot);
      GammaDiscreteRateDistribution g(4, 0.7);
      DRHomogeneousTreeLikelihood tl(*t, s, &lg, &g);
      run("PN_LG08", tl, 1e-6);
    }
  } catch (std::exception& e) {
    cerr << e.what() << endl;
    return 1;
  }
  return 0;
}
