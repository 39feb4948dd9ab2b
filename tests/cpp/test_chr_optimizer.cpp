// The multi-start flow of ChromosomeNumberMng::runChromEvol (App/ChromosomeNumberMng.cpp:264-288) through the shim on the device:
// ChromosomeNumberOptimizer::optimize over several starting points (all points' Brent probes in one device call per step), then
// the joint ML and the marginal reconstruction at the best point.  tests/test_cpp_shim.py re-evaluates the printed parameter points
// with the oracle.
#include <cmath>
#include <cstdio>
#include <iostream>
#include <memory>

#include "../../bpp_phyl_b200/host/bppgpu_shim.hpp"

using namespace bppshim;
using namespace std;

int main() {
  try {
    ChromosomeAlphabet chr(1, 30);
    unique_ptr<Tree> tree(TreeTemplateTools::parenthesisToTree("(((a:0.3,b:0.2):0.4,c:0.5):0.1,(d:0.3,e:0.6):0.2);"));
    VectorSiteContainer sites(&chr);
    sites.addSequence(BasicSequence("a", "7", &chr));
    sites.addSequence(BasicSequence("b", "8", &chr));
    sites.addSequence(BasicSequence("c", "14", &chr));
    sites.addSequence(BasicSequence("d", "9", &chr));
    sites.addSequence(BasicSequence("e", "X", &chr));
    const double starts[6][4] = {{0.7, 0.4, 0.2, 0.1}, {1.1, 0.4, 0.2, 0.05}, {0.2, 1.3, 0.6, 0.3}, {2.0, 2.0, 0.01, 0.4}, {0.05, 0.05, 0.9, 0.02}, {5.0, 0.5, 0.1, 1.0}};
    vector<unique_ptr<ChromosomeSubstitutionModel> > own;
    vector<ChromosomeSubstitutionModel*> models;
    for (int k = 0; k < 6; ++k) {
      own.emplace_back(new ChromosomeSubstitutionModel(&chr, starts[k][0], starts[k][1], starts[k][2], starts[k][3]));
      models.push_back(own.back().get());
    }
    ConstantRateDistribution cst;
    ChromosomeNumberOptimizer opt(*tree, sites, models, &cst, /*weightedRootFreq=*/true);
    for (size_t k = 0; k < 6; ++k) {
      printf("OPT_START_%zu %.15f\n", k, opt.getValue(k));
      printf("OPT_START_NONSINGULAR_%zu %d\n", k, (int)models[k]->isNonSingular());
    }
    // searched on [1e-3, 3] (the reference: (0, 100]) to keep the test short
    opt.optimize({6, 3, 1}, {0, 2, 3}, 1e-3, 1e-3, 3.0);
    const vector<size_t>& order = opt.getPointOrder();
    for (size_t r = 0; r < 6; ++r) printf("OPT_ORDER_%zu %zu\n", r, order[r]);
    for (size_t k = 0; k < 6; ++k) {
      printf("OPT_FINAL_%zu %.15f\n", k, opt.getValue(k));
      const char* names[4] = {"gain", "loss", "dupl", "demi"};
      for (int j = 0; j < 4; ++j) printf("OPT_PARAM_%zu_%s %.17g\n", k, names[j], opt.getModel(k)->getParameterValue(names[j]));
    }
    printf("OPT_BEST %.15f\n", opt.getBestValue());
    printf("OPT_BATCH_EVALS %u\n", opt.getNumberOfBatchEvaluations());
    printf("OPT_POINT_EVALS %ld\n", opt.getLikelihoods().getNumberOfLikelihoodCalculations());

    // runChromEvol's next steps on the best point: a single likelihood on the optimised model, joint ML + marginal reconstruction
    DRNonHomogeneousTreeLikelihood tl(*tree, sites, true, false, opt.getBestModel(), &cst);
    tl.initialize();
    printf("OPT_BEST_SINGLE %.15f\n", tl.getValue());
    printf("OPT_BEST_NONSINGULAR %d\n", (int)opt.getBestModel()->isNonSingular());
    MLAncestralStateReconstruction ml(&tl, opt.getBestModel(), tl.getRootFrequencies());
    ml.computeJointLikelihood();
    for (auto& kv : ml.getAllAncestralStates()) printf("OPT_ML_%d %zu\n", kv.first, kv.second[0]);
    printf("OPT_ML_BEST_LNL %.15g\n", ml.getBestJointLogLikelihoodPerSite()[0]);
    MarginalNonRevAncestralStateReconstruction asr(&tl);
    asr.computePosteriorProbabilitiesOfNodesForEachStatePerSite();
    for (auto& kv : asr.getAllAncestralStates()) {
      printf("OPT_MARG_%d %zu\n", kv.first, kv.second[0]);
      printf("OPT_MARGP_%d %.15g\n", kv.first, (*asr.getPosteriorProbForAllNodesAndStatesPerSite())[kv.first][0][kv.second[0]]);
    }
  } catch (std::exception& e) {
    cerr << e.what() << endl;
    return 1;
  }
  return 0;
}
