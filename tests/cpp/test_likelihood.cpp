// Mirrors the reference's own likelihood tests through the C++ host shim (GPU required):
//   test/test_likelihood.cpp:90-136        T92(kappa=3)+Gamma4 on 4 taxa x 17 sites, R and DR classes, -lnL = 85.030942031997312824,
//                                          first derivatives of the two classes agree to 1e-6
//   test/test_likelihood_clock.cpp:99-115  rooted tree kept rooted (checkRooted=false), constant rate: 94.3957
// plus derivative-vs-finite-difference checks and the error conventions of SURVEY.md 8b.  Exit code 0 = success
// (the reference's CTest convention).  Extra cases print "name value" lines that tests/test_cpp_shim.py compares with the oracle.
#include <cmath>
#include <cstdio>
#include <iostream>
#include <memory>

#include "../../bpp_phyl_b200/host/bppgpu_shim.hpp"

using namespace bppshim;
using namespace std;

static int fails = 0;
#define EXPECT(cond, msg)                                     \
  do {                                                        \
    if (!(cond)) { cerr << "FAILED: " << msg << endl; ++fails; } \
  } while (0)

template <class LIK>
void fitModel(SubstitutionModel* model, DiscreteDistribution* rdist, const Tree& tree, const VectorSiteContainer& sites, double initialValue,
              const char* label) {
  LIK tl(tree, sites, model, rdist);
  tl.initialize();
  printf("%s %.15f\n", label, tl.getValue());
  EXPECT(fabs(tl.getValue() - initialValue) < 1e-9, "Incorrect initial value (" << label << "): " << tl.getValue());
}

int main() {
  try {
    unique_ptr<Tree> tree(TreeTemplateTools::parenthesisToTree("((A:0.01, B:0.02):0.03,C:0.01,D:0.1);"));
    const DNA* alphabet = &AlphabetTools::DNA_ALPHABET();
    VectorSiteContainer sites(alphabet);
    sites.addSequence(BasicSequence("A", "AAATGGCTGTGCACGTC", alphabet));
    sites.addSequence(BasicSequence("B", "GACTGGATCTGCACGTC", alphabet));
    sites.addSequence(BasicSequence("C", "CTCTGGATGTGCACGTG", alphabet));
    sites.addSequence(BasicSequence("D", "AAATGGCGGTGCGCCTA", alphabet));

    unique_ptr<SubstitutionModel> model(new T92(alphabet, 3.));
    unique_ptr<DiscreteDistribution> rdist(new GammaDiscreteRateDistribution(4, 1.0));
    cout << "Testing Single Tree Traversal likelihood class..." << endl;
    fitModel<RHomogeneousTreeLikelihood>(model.get(), rdist.get(), *tree, sites, 85.030942031997312824, "R_T92_G4");
    cout << "Testing Double Tree Traversal likelihood class..." << endl;
    fitModel<DRHomogeneousTreeLikelihood>(model.get(), rdist.get(), *tree, sites, 85.030942031997312824, "DR_T92_G4");

    // Let's compare the derivatives (test_likelihood.cpp:124-135) and check them against finite differences
    RHomogeneousTreeLikelihood tlsr(*tree, sites, model.get(), rdist.get());
    tlsr.initialize();
    DRHomogeneousTreeLikelihood tldr(*tree, sites, model.get(), rdist.get());
    tldr.initialize();
    ParameterList params = tlsr.getBranchLengthsParameters();
    EXPECT(params.size() == 5, "an unrooted 4-taxon tree has 5 branch-length parameters, got " << params.size());
    for (const Parameter& p : params) {
      const double d1sr = tlsr.getFirstOrderDerivative(p.name), d1dr = tldr.getFirstOrderDerivative(p.name);
      const double d2dr = tldr.getSecondOrderDerivative(p.name);
      printf("%s\t%.12g\t%.12g\t%.12g\n", p.name.c_str(), d1sr, d1dr, d2dr);
      EXPECT(fabs(d1sr - d1dr) <= 0.000001, "R and DR first derivatives differ for " << p.name);
      const double h = 1e-6, f0 = tldr.getValue();
      tldr.setParameterValue(p.name, p.value + h);
      const double fp = tldr.getValue();
      tldr.setParameterValue(p.name, p.value - h);
      const double fm = tldr.getValue();
      tldr.setParameterValue(p.name, p.value);
      EXPECT(fabs((fp - fm) / (2 * h) - d1dr) < 1e-4 * max(1.0, fabs(d1dr)), "first derivative vs finite difference, " << p.name);
      EXPECT(fabs(tldr.getValue() - f0) < 1e-12, "value restored");
      (void)d2dr;
    }
    // per-site accessors: sum over sites of log L_site = log L, duplicated columns share a pattern
    double s = 0;
    for (size_t i = 0; i < tldr.getNumberOfSites(); ++i) s += tldr.getLogLikelihoodForASite(i);
    EXPECT(fabs(s - tldr.getLogLikelihood()) < 1e-10, "sum of per-site log-likelihoods");
    EXPECT(tldr.getNumberOfSites() == 17 && tldr.getNumberOfDistinctSites() == 12, "17 sites, 12 patterns");

    // getPij_t interface (Interface 1): rows sum to one, P(0) = I, dP rows sum to zero
    const RowMatrix<double>& P = model->getPij_t(0.1);
    for (size_t i = 0; i < 4; ++i) {
      double r = 0;
      for (size_t j = 0; j < 4; ++j) r += P(i, j);
      EXPECT(fabs(r - 1.0) < 1e-14, "row sum of P(t)");
    }
    EXPECT(fabs(model->Pij_t(2, 2, 0.0) - 1.0) < 1e-15 && fabs(model->Pij_t(2, 1, 0.0)) < 1e-15, "P(0) = I");

    // clock test: rooted tree, kept rooted, constant rate (test_likelihood_clock.cpp:99-115)
    {
      unique_ptr<Tree> t2(TreeTemplateTools::parenthesisToTree("(((A:0.01, B:0.01):0.02,C:0.03):0.01,D:0.04);"));
      VectorSiteContainer s2(alphabet);
      s2.addSequence(BasicSequence("A", "AAATGGCTGTGCACGTC", alphabet));
      s2.addSequence(BasicSequence("B", "AACTGGATCTGCATGTC", alphabet));
      s2.addSequence(BasicSequence("C", "ATCTGGACGTGCACGTG", alphabet));
      s2.addSequence(BasicSequence("D", "CAACGGGAGTGCGCCTA", alphabet));
      ConstantRateDistribution cst;
      RHomogeneousTreeLikelihood tl(*t2, s2, model.get(), &cst, false, false);
      tl.initialize();
      printf("CLOCK_T92_CONST %.15f\n", tl.getValue());
      EXPECT(fabs(tl.getValue() - 94.3957) < 1e-4, "Incorrect initial value (clock): " << tl.getValue());
    }

    // error conventions
    {
      RHomogeneousTreeLikelihood tl(*tree, sites, model.get(), rdist.get());
      bool thrown = false;
      try { tl.getValue(); } catch (Exception&) { thrown = true; }
      EXPECT(thrown, "getValue() before initialize() must throw");
      tl.initialize();
      thrown = false;
      try { tl.initialize(); } catch (Exception&) { thrown = true; }
      EXPECT(thrown, "second initialize() must throw");
      thrown = false;
      try { tl.getFirstOrderDerivative("T92.kappa"); } catch (Exception&) { thrown = true; }
      EXPECT(thrown, "derivative w.r.t. a model parameter must throw");
      thrown = false;
      try { tl.getFirstOrderDerivative("BrLen99"); } catch (ParameterNotFoundException&) { thrown = true; }
      EXPECT(thrown, "unknown parameter must throw ParameterNotFoundException");
      // a model parameter change re-evaluates everything (fireParameterChanged)
      const double before = tl.getValue();
      tl.setParameterValue("T92.kappa", 2.0);
      EXPECT(fabs(tl.getValue() - before) > 1e-3, "kappa change must move the likelihood");
      tl.setParameterValue("T92.kappa", 3.0);
      EXPECT(fabs(tl.getValue() - before) < 1e-10, "kappa restored");
      model->setParameterValue("kappa", 3.0);
    }

    // other state spaces: values printed for the oracle comparison in tests/test_cpp_shim.py
    {
      const ProteicAlphabet* prot = &AlphabetTools::PROTEIN_ALPHABET();
      unique_ptr<Tree> t3(TreeTemplateTools::parenthesisToTree("((a:0.1,b:0.2):0.05,(c:0.3,d:0.02):0.07,e:0.15);"));
      VectorSiteContainer s3(prot);
      s3.addSequence(BasicSequence("a", "ARNDCQEGHILKMFPSTWYVAAX", prot));
      s3.addSequence(BasicSequence("b", "ARNDCQEGHILKMFPSTWYVLK-", prot));
      s3.addSequence(BasicSequence("c", "ARNECQDGHLIKMFPTSWYVAKB", prot));
      s3.addSequence(BasicSequence("d", "GRNDCQEGHILRMYPSTWFVAAZ", prot));
      s3.addSequence(BasicSequence("e", "ARNDCQEGHVLKMFPSTWYIVAA", prot));
      LG08 lg(prot);
      GammaDiscreteRateDistribution g4(4, 0.7);
      DRHomogeneousTreeLikelihood tl(*t3, s3, &lg, &g4);
      tl.initialize();
      printf("LG08_G4 %.15f\n", tl.getValue());
      for (const Parameter& p : tl.getBranchLengthsParameters())
        printf("LG08_G4_d %s %.12g %.12g\n", p.name.c_str(), tl.getFirstOrderDerivative(p.name), tl.getSecondOrderDerivative(p.name));
    }
    {
      const CodonAlphabet* cod = &AlphabetTools::CODON_ALPHABET();
      unique_ptr<Tree> t4(TreeTemplateTools::parenthesisToTree("((a:0.1,b:0.2):0.05,c:0.3,d:0.02);"));
      VectorSiteContainer s4(cod);
      s4.addSequence(BasicSequence("a", "ATGGCTAAATTTGGGCCC", cod));
      s4.addSequence(BasicSequence("b", "ATGGCCAAATTCGGGCCA", cod));
      s4.addSequence(BasicSequence("c", "ATGGCTAAGTTTGGACCC", cod));
      s4.addSequence(BasicSequence("d", "ATGTCTAAATTTGGGCCC", cod));
      YN98 yn(cod, 2.0, 0.3);
      ConstantRateDistribution cst;
      DRHomogeneousTreeLikelihood tl(*t4, s4, &yn, &cst);
      tl.initialize();
      printf("YN98_CONST %.15f\n", tl.getValue());
    }
    {
      // consumers of the DR arrays: posterior probabilities at every node sum to one, computeLikelihoodAtNode integrates to the
      // site likelihood at every node (DRTreeLikelihoodTools / MarginalAncestralStateReconstruction)
      const DNA dna2;
      unique_ptr<Tree> t6(TreeTemplateTools::parenthesisToTree("((A:0.01, B:0.02):0.03,C:0.01,D:0.1);"));
      VectorSiteContainer s6(&dna2);
      s6.addSequence(BasicSequence("A", "AAATGGCTGTGCACGTC", &dna2));
      s6.addSequence(BasicSequence("B", "GACTGGATCTGCACGTC", &dna2));
      s6.addSequence(BasicSequence("C", "CTCTGGATGTGCACGTG", &dna2));
      s6.addSequence(BasicSequence("D", "AAATGGCGGTGCGCCTA", &dna2));
      T92 m6(&dna2, 3.);
      GammaDiscreteRateDistribution g6(4, 1.0);
      DRHomogeneousTreeLikelihood tl(*t6, s6, &m6, &g6);
      tl.initialize();
      double worst = 0, worstL = 0;
      for (int nid = 0; nid < 6; ++nid) {
        VVVdouble post = tl.getPosteriorProbabilitiesForEachStateForEachRate(nid);
        VVVdouble full;
        tl.computeLikelihoodAtNode(nid, full);
        for (size_t i = 0; i < post.size(); ++i) {
          double sp = 0, sl = 0;
          for (size_t c = 0; c < post[i].size(); ++c)
            for (size_t x = 0; x < post[i][c].size(); ++x) { sp += post[i][c][x]; sl += full[i][c][x] * g6.getProbability(c); }
          worst = max(worst, fabs(sp - 1.0));
          size_t site = 0;
          while (tl.getSiteIndex(site) != i) ++site;
          worstL = max(worstL, fabs(log(sl) - tl.getLogLikelihoodForASite(site)));
        }
      }
      // MarginalAncestralStateReconstruction::getAncestralStatesForNode: probs = sum_c computeLikelihoodAtNode r_c / l_i (.cpp:70-83),
      // here from another device kernel (the marginal posterior table) than computeLikelihoodAtNode
      MarginalAncestralStateReconstruction masr(&tl);
      double masrErr = 0;
      for (int nid : {2, 4, 5}) {   // an internal node, a leaf, the root
        VVdouble probs;
        const vector<size_t> st = masr.getAncestralStatesForNode(nid, probs);
        VVVdouble full;
        tl.computeLikelihoodAtNode(nid, full);
        for (size_t i = 0; i < probs.size(); ++i) {
          size_t site = 0;
          while (tl.getSiteIndex(site) != i) ++site;
          size_t arg = 0;
          for (size_t x = 0; x < 4; ++x) {
            double v = 0;
            for (size_t c = 0; c < 4; ++c) v += full[i][c][x] * g6.getProbability(c);
            v /= tl.getLikelihoodForASite(site);
            masrErr = max(masrErr, fabs(v - probs[i][x]));
            if (probs[i][x] > probs[i][arg]) arg = x;
          }
          if (arg != st[i]) masrErr = 1;
        }
      }
      VVdouble leafProbs;
      const vector<size_t> leafStates = masr.getAncestralStatesForNode(0, leafProbs);   // leaf A: its own characters
      const string seqA = "AAATGGCTGTGCACGTC";
      for (size_t site = 0; site < seqA.size(); ++site)
        if (string("ACGT")[leafStates[tl.getSiteIndex(site)]] != seqA[site]) masrErr = 1;
      // DiscreteRatesAcrossSitesTreeLikelihood accessors (posterior rates per site etc.) from the root arrays of the device
      {
        const Vdouble pr = tl.getPosteriorRateOfEachSite();
        const vector<size_t> mc = tl.getRateClassWithMaxPostProbOfEachSite();
        const VVdouble pb = tl.getPosteriorProbabilitiesOfEachRate();
        double drasErr = 0;
        for (size_t i = 0; i < pr.size(); ++i) {
          printf("POSTRATE_%zu %.15g\n", i, pr[i]);
          printf("MAXCLASS_%zu %zu\n", i, mc[i]);
          double sp = 0, sl = 0, ss = 0;
          for (size_t c = 0; c < 4; ++c) {
            sp += pb[i][c];
            sl += tl.getLikelihoodForASiteForARateClass(i, c) * g6.getProbability(c);
          }
          for (int x = 0; x < 4; ++x) ss += tl.getLikelihoodForASiteForAState(i, x) * tl.getRootFrequencies()[x];
          drasErr = max(drasErr, fabs(sp - 1.0));
          drasErr = max(drasErr, fabs(sl / tl.getLikelihoodForASite(i) - 1.0));   // sum_c p_c L[i][c] = L[i]
          drasErr = max(drasErr, fabs(ss / tl.getLikelihoodForASite(i) - 1.0));   // sum_x pi_x sum_c p_c L[i][c][x] = L[i]
        }
        printf("DRAS_ERR %.3e\n", drasErr);
        if (drasErr > 1e-12) { cerr << "DiscreteRatesAcrossSites accessor identities failed" << endl; fails++; }
      }
      printf("MASR_ERR %.3e\n", masrErr);
      printf("MASR_NNODES %zu\n", masr.getAllAncestralStates().size());
      if (masrErr > 1e-12) { cerr << "MarginalAncestralStateReconstruction check failed" << endl; fails++; }
      printf("POSTERIOR_SUM_ERR %.3e\n", worst);
      printf("ATNODE_LNL_ERR %.3e\n", worstL);
      vector<size_t> anc = tl.getAncestralStatesForNode(5);
      printf("ANCESTRAL_ROOT_SITE0 %zu\n", anc[0]);
      if (worst > 1e-12 || worstL > 1e-10) { cerr << "posterior / computeLikelihoodAtNode checks failed" << endl; fails++; }
    }
    {
      // mixture of sub-models (RHomogeneousMixedTreeLikelihood): three T92 with different kappa, probabilities .2 / .5 / .3
      const DNA dna3;
      unique_ptr<Tree> t7(TreeTemplateTools::parenthesisToTree("((A:0.01, B:0.02):0.03,C:0.01,D:0.1);"));
      VectorSiteContainer s7(&dna3);
      s7.addSequence(BasicSequence("A", "AAATGGCTGTGCACGTC", &dna3));
      s7.addSequence(BasicSequence("B", "GACTGGATCTGCACGTC", &dna3));
      s7.addSequence(BasicSequence("C", "CTCTGGATGTGCACGTG", &dna3));
      s7.addSequence(BasicSequence("D", "AAATGGCGGTGCGCCTA", &dna3));
      T92 ma(&dna3, 1.), mb(&dna3, 3.), mc(&dna3, 8.);
      GammaDiscreteRateDistribution g7(4, 1.0);
      vector<SubstitutionModel*> subs;
      subs.push_back(&ma); subs.push_back(&mb); subs.push_back(&mc);
      Vdouble pr(3);
      pr[0] = 0.2; pr[1] = 0.5; pr[2] = 0.3;
      RHomogeneousMixedTreeLikelihood mix(*t7, s7, subs, pr, &g7);
      mix.initialize();
      printf("MIXED_T92_G4 %.15f\n", mix.getValue());
      // a degenerate mixture is the plain likelihood
      Vdouble one(3, 0.0);
      one[1] = 1.0;
      mix.setProbabilities(one);
      printf("MIXED_DEGENERATE %.15f\n", mix.getValue());
      if (fabs(mix.getValue() - 85.030942031997312824) > 1e-9) { cerr << "degenerate mixture != golden value" << endl; fails++; }
    }
    {
      // BrLenRoot / RootPosition (reparametrizeRoot, test_likelihood_nh.cpp:57-58): derivatives against central differences
      const DNA dna4;
      unique_ptr<Tree> t8(TreeTemplateTools::parenthesisToTree("(((A:0.01, B:0.01):0.02,C:0.03):0.01,D:0.04);"));
      VectorSiteContainer s8(&dna4);
      s8.addSequence(BasicSequence("A", "AAATGGCTGTGCACGTC", &dna4));
      s8.addSequence(BasicSequence("B", "AACTGGATCTGCATGTC", &dna4));
      s8.addSequence(BasicSequence("C", "ATCTGGACGTGCACGTG", &dna4));
      s8.addSequence(BasicSequence("D", "CAACGGGAGTGCGCCTA", &dna4));
      T92 m8(&dna4, 3.);
      GammaDiscreteRateDistribution g8(4, 1.0);
      Vdouble rf(4);
      rf[0] = 0.1; rf[1] = 0.4; rf[2] = 0.3; rf[3] = 0.2;   // non-stationary root: the root position matters
      DRNonHomogeneousTreeLikelihood tl(*t8, s8, false, true, &m8, &g8, &rf, false, 0, true);
      tl.initialize();
      const char* names[2] = {"BrLenRoot", "RootPosition"};
      for (int k = 0; k < 2; ++k) {
        const double v0 = tl.getParameterValue(names[k]), h = 1e-5 * (k == 0 ? 1.0 : 1.0);
        const double d1 = tl.getFirstOrderDerivative(names[k]), d2 = tl.getSecondOrderDerivative(names[k]);
        const double f0 = tl.getValue();
        tl.setParameterValue(names[k], v0 + h);
        const double fp = tl.getValue();
        tl.setParameterValue(names[k], v0 - h);
        const double fm = tl.getValue();
        tl.setParameterValue(names[k], v0);
        const double fd1 = (fp - fm) / (2 * h), fd2 = (fp - 2 * f0 + fm) / (h * h);
        printf("REPARAM_%s %.12f %.12f %.9f %.9f\n", names[k], d1, fd1, d2, fd2);
        if (fabs(d1 - fd1) > 1e-5 * max(1.0, fabs(d1)) || fabs(d2 - fd2) > 2e-3 * max(1.0, fabs(d2))) {
          cerr << "BrLenRoot / RootPosition derivative differs from finite differences" << endl;
          fails++;
        }
      }
    }
    {
      // test/test_likelihood_clock.cpp:80-93, 113-121: the clock class starts at the unconstrained value (the tree is ultrametric;
      // the reference's 92.3295 is that start with the model parameters its first fit left behind) and, at the clock-constrained
      // optimum, must give the reference's final value 71.2657 (tolerance 0.001 there; argmin from tests/golden/clock_optimum.json)
      const DNA dna5;
      unique_ptr<Tree> t9(TreeTemplateTools::parenthesisToTree("(((A:0.01, B:0.01):0.02,C:0.03):0.01,D:0.04);"));
      VectorSiteContainer s9(&dna5);
      s9.addSequence(BasicSequence("A", "AAATGGCTGTGCACGTC", &dna5));
      s9.addSequence(BasicSequence("B", "AACTGGATCTGCATGTC", &dna5));
      s9.addSequence(BasicSequence("C", "ATCTGGACGTGCACGTG", &dna5));
      s9.addSequence(BasicSequence("D", "CAACGGGAGTGCGCCTA", &dna5));
      T92 m9(&dna5, 3.);
      ConstantRateDistribution c9;
      RHomogeneousClockTreeLikelihood tl(*t9, s9, &m9, &c9);
      tl.initialize();
      printf("CLOCK_INIT %.15f\n", tl.getValue());
      printf("CLOCK_TOTALHEIGHT %.15f\n", tl.getParameterValue("TotalHeight"));
      printf("CLOCK_NPARAMS %zu\n", tl.getBranchLengthsParameters().size());
      tl.setParametersValues({{"TotalHeight", 0.448120364718}, {"HeightP2", 0.818792488373}, {"HeightP4", 0.390971377099},
                              {"T92.kappa", 0.996348825463}, {"T92.theta", 0.579945474376}});
      printf("CLOCK_OPTIMUM %.15f\n", tl.getValue());
      if (fabs(tl.getValue() - 71.2657) > 0.001) { cerr << "Incorrect final value (clock)." << endl; fails++; }
      // a crude clock-constrained descent from the reference's starting point must also come down to it: cyclic golden-section
      // line searches over the five parameters (the surface is smooth; this is not the reference's optimiser)
      m9.setParameterValue("kappa", 3.0);   // the model object is shared with tl, which left it at the optimum
      m9.setParameterValue("theta", 0.5);
      RHomogeneousClockTreeLikelihood tl2(*t9, s9, &m9, &c9);
      tl2.initialize();
      const char* names[5] = {"TotalHeight", "HeightP2", "HeightP4", "T92.kappa", "T92.theta"};
      double x[5] = {0.04, 1.0 / 3.0, 0.75, 3.0, 0.5};
      const double lo[5] = {1e-6, 1e-6, 1e-6, 1e-3, 1e-3}, hi[5] = {2.0, 1 - 1e-6, 1 - 1e-6, 20.0, 0.999};
      double best = tl2.getValue();
      for (int sweep = 0; sweep < 60; ++sweep) {
        const double before = best;
        for (int k = 0; k < 5; ++k) {
          double a = lo[k], b = hi[k];
          const double g = 0.6180339887498949;
          double c1 = b - g * (b - a), c2 = a + g * (b - a);
          tl2.setParameterValue(names[k], c1); double f1 = tl2.getValue();
          tl2.setParameterValue(names[k], c2); double f2 = tl2.getValue();
          for (int it = 0; it < 60; ++it) {
            if (f1 < f2) { b = c2; c2 = c1; f2 = f1; c1 = b - g * (b - a); tl2.setParameterValue(names[k], c1); f1 = tl2.getValue(); }
            else { a = c1; c1 = c2; f1 = f2; c2 = a + g * (b - a); tl2.setParameterValue(names[k], c2); f2 = tl2.getValue(); }
          }
          const double xm = 0.5 * (a + b);
          tl2.setParameterValue(names[k], xm);
          if (tl2.getValue() <= best) { best = tl2.getValue(); x[k] = xm; }
          else tl2.setParameterValue(names[k], x[k]);
        }
        if (before - best < 1e-9) break;
      }
      printf("CLOCK_DESCENT %.15f\n", best);
      if (fabs(best - 71.2657) > 0.001) { cerr << "clock-constrained descent did not reach the reference optimum" << endl; fails++; }
    }
    {
      // test/test_likelihood_nh.cpp:73-110: one T92 per branch (kappa shared, theta free), GC root frequencies, Gamma(4, 1);
      // the statistical recovery loop of that test is the optimiser's business, the likelihood of the set is checked here
      const DNA dna4;
      unique_ptr<Tree> t8(TreeTemplateTools::parenthesisToTree("(((A:0.1, B:0.2):0.3,C:0.1):0.2,(D:0.3,(E:0.2,F:0.05):0.1):0.1);"));
      VectorSiteContainer s8(&dna4);
      s8.addSequence(BasicSequence("A", "ATGTTATCCCGTCGAATCATATGGAATCGTCTAGAACTCA", &dna4));
      s8.addSequence(BasicSequence("B", "ATGGTATCTCGCCTAATCATGTGGCATCGTCAAAAAATCA", &dna4));
      s8.addSequence(BasicSequence("C", "TTGGTGTGTCCCTTAATCGTGTGGTATCGTCCGGATATAG", &dna4));
      s8.addSequence(BasicSequence("D", "ATGGAATCTCCCCTATTCAAGTGGTAACGTCTAGAAATAA", &dna4));
      s8.addSequence(BasicSequence("E", "CTGGTATCTCCCATTATCATGTGCTATAGTCGAAAAACAA", &dna4));
      s8.addSequence(BasicSequence("F", "ATGGTTTCTCCCCTAATAGTTCGGCAACGTCAAGACATCA", &dna4));
      FrequencySet* rootFreqs = new GCFrequencySet(&dna4);
      SubstitutionModel* model = new T92(&dna4, 3.);
      unique_ptr<SubstitutionModelSet> modelSet(SubstitutionModelSetTools::createNonHomogeneousModelSet(model, rootFreqs, t8.get(), {"T92.kappa"}));
      const size_t nmodels = modelSet->getNumberOfModels();
      for (size_t i = 0; i < nmodels; ++i) modelSet->setParameterValue("T92.theta_" + to_string(i + 1), 0.15 + 0.07 * (double)i);
      GammaDiscreteRateDistribution g8(4, 1.0);
      unique_ptr<SubstitutionModelSet> modelSet2(modelSet->clone());
      DRNonHomogeneousTreeLikelihood tl(*t8, s8, modelSet.get(), &g8, false, false);
      tl.initialize();
      RNonHomogeneousTreeLikelihood tlR(*t8, s8, modelSet2.get(), &g8, false, true, false);
      tlR.initialize();
      printf("NH_NMODELS %zu\n", nmodels);
      printf("NH_DR_T92_G4 %.15f\n", tl.getValue());
      printf("NH_R_T92_G4 %.15f\n", tlR.getValue());
      printf("NH_NPARAMS %zu\n", tl.getSubstitutionModelParameters().size());
      for (const Parameter& p : tl.getBranchLengthsParameters())
        printf("NH_T92_G4_d %s %.12g %.12g\n", p.name.c_str(), tl.getFirstOrderDerivative(p.name), tl.getSecondOrderDerivative(p.name));
      if (fabs(tl.getValue() - tlR.getValue()) > 1e-9) { cerr << "R and DR non-homogeneous likelihoods differ" << endl; fails++; }
      // move the shared kappa (all ten models follow), one branch's theta and the root GC content
      tl.setParametersValues({{"T92.kappa_1", 2.0}, {"T92.theta_3", 0.6}, {"GC.theta", 0.3}});
      printf("NH_DR_T92_G4_MOVED %.15f\n", tl.getValue());
      printf("NH_KAPPA_7 %.15f\n", modelSet->getParameterValue("T92.kappa_7"));
    }
    {
      ChromosomeAlphabet chr(1, 30);
      unique_ptr<Tree> t5(TreeTemplateTools::parenthesisToTree("(((a:0.3,b:0.2):0.4,c:0.5):0.1,(d:0.3,e:0.6):0.2);"));
      VectorSiteContainer s5(&chr);
      s5.addSequence(BasicSequence("a", "7", &chr));
      s5.addSequence(BasicSequence("b", "8", &chr));
      s5.addSequence(BasicSequence("c", "14", &chr));
      s5.addSequence(BasicSequence("d", "9", &chr));
      s5.addSequence(BasicSequence("e", "X", &chr));
      ChromosomeSubstitutionModel cm(&chr, 0.7, 0.4, 0.2, ChromosomeSubstitutionModel::DemiEqualDupl);
      ConstantRateDistribution cst;
      DRNonHomogeneousTreeLikelihood tl(*t5, s5, true, false, &cm, &cst);
      tl.initialize();
      printf("CHR_WEIGHTED %.15f\n", tl.getValue());
      {
        // marginal reconstruction for non-reversible models (fork: MarginalNonRevAncestralStateReconstruction), as
        // ChromosomeNumberMng::runChromEvol does after the optimisation
        MarginalNonRevAncestralStateReconstruction asr(&tl);
        asr.computePosteriorProbabilitiesOfNodesForEachStatePerSite();
        auto* post = asr.getPosteriorProbForAllNodesAndStatesPerSite();
        auto joint = asr.getAllJointFatherNodeProbabilities();
        const map<int, vector<size_t> > anc = asr.getAllAncestralStates();
        double sumErr = 0, jointErr = 0, fatherErr = 0;
        const int rootId = (int)post->size() - 1;
        for (auto& kv : *post) {
          const vector<double>& p = kv.second[0];
          double s = 0;
          for (double v : p) s += v;
          sumErr = max(sumErr, fabs(s - 1.0));
          printf("CHR_ANC_%d %zu\n", kv.first, anc.at(kv.first)[0]);
          printf("CHR_POSTMAX_%d %.15g\n", kv.first, p[anc.at(kv.first)[0]]);
          if (kv.first == rootId) continue;
          const VVdouble& j = joint[kv.first][0];
          const int father = tl.getTree().getNode(kv.first)->getFather()->getId();
          const vector<double>& pf = (*post)[father][0];
          for (size_t x = 0; x < p.size(); ++x) {
            double r = 0;
            for (size_t y = 0; y < p.size(); ++y) r += j[x][y];
            jointErr = max(jointErr, fabs(r - p[x]));
          }
          for (size_t y = 0; y < p.size(); ++y) {
            double cs = 0;
            for (size_t x = 0; x < p.size(); ++x) cs += j[x][y];
            fatherErr = max(fatherErr, fabs(cs - pf[y]));
          }
        }
        // joint ML reconstruction (fork: MLAncestralStateReconstruction), also called by runChromEvol
        MLAncestralStateReconstruction mlasr(&tl, &cm, tl.getRootFrequencies());
        mlasr.computeJointLikelihood();
        const map<int, vector<size_t> > mlStates = mlasr.getAllAncestralStates();
        for (auto& kv : mlStates) printf("CHR_ML_%d %zu\n", kv.first, kv.second[0]);
        printf("CHR_ML_BEST %.15g\n", mlasr.getBestJointLogLikelihoodPerSite()[0]);
        const vector<double> rp = asr.getRootPosteriorProb();
        double rootErr = 0;
        for (size_t x = 0; x < rp.size(); ++x) rootErr = max(rootErr, fabs(rp[x] - (*post)[rootId][0][x]));
        printf("CHR_MARG_SUM_ERR %.3e\n", sumErr);
        printf("CHR_MARG_JOINT_ERR %.3e\n", jointErr);
        printf("CHR_MARG_FATHER_ERR %.3e\n", fatherErr);
        printf("CHR_MARG_ROOT_ERR %.3e\n", rootErr);
        if (sumErr > 1e-10 || jointErr > 1e-12 || fatherErr > 1e-10 || rootErr > 1e-15) {
          cerr << "MarginalNonRevAncestralStateReconstruction consistency checks failed" << endl;
          fails++;
        }
      }
      cm.setParameterValue("Chromosome.gain", 1.1);
      // the batched front-end: ChromosomeNumberOptimizer's vector of likelihoods (one starting point each) as one device object
      const double pts[5][4] = {{0.7, 0.4, 0.2, 0.1}, {1.1, 0.4, 0.2, 0.05}, {0.2, 1.3, 0.6, 0.3}, {2.0, 2.0, 0.01, 0.4}, {0.05, 0.05, 0.9, 0.0}};
      vector<unique_ptr<ChromosomeSubstitutionModel> > owned;
      vector<SubstitutionModel*> models;
      for (int k = 0; k < 5; ++k) {
        owned.emplace_back(new ChromosomeSubstitutionModel(&chr, pts[k][0], pts[k][1], pts[k][2], pts[k][3]));
        models.push_back(owned.back().get());
      }
      LikelihoodPointBatch batch(*t5, s5, true, models, &cst);
      batch.initialize();
      double maxrel = 0;
      for (int k = 0; k < 5; ++k) {
        DRNonHomogeneousTreeLikelihood one(*t5, s5, true, false, models[k], &cst);
        one.initialize();
        const double rel = fabs(batch.getValue(k) - one.getValue()) / fabs(one.getValue());
        if (rel > maxrel) maxrel = rel;
        printf("CHR_BATCH_%d %.15f\n", k, batch.getValue(k));
      }
      printf("CHR_BATCH_MAXREL %.3e\n", maxrel);
      // a Brent-style probe of one point: change a parameter of point 2, re-evaluate everything in one call
      owned[2]->setParameterValue("Chromosome.loss", 0.9);
      batch.modelChanged(2);
      DRNonHomogeneousTreeLikelihood probe(*t5, s5, true, false, models[2], &cst);
      probe.initialize();
      printf("CHR_BATCH_PROBE_REL %.3e\n", fabs(batch.getValue(2) - probe.getValue()) / fabs(probe.getValue()));
      printf("CHR_BATCH_BEST %zu\n", batch.getBestPoint());
      if (maxrel > 1e-12) { cerr << "batched points differ from single-point likelihoods" << endl; fails++; }
    }
  } catch (exception& ex) {
    cerr << "EXCEPTION: " << ex.what() << endl;
    return 1;
  }
  if (fails) { cerr << fails << " check(s) failed" << endl; return 1; }
  cout << "OK" << endl;
  return 0;
}
