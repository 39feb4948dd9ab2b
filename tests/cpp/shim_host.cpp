// Host-side (no GPU) checks of the C++ shim: prints one JSON document that tests/test_cpp_shim.py compares with the oracle.
#include <cstdio>
#include <iostream>
#include <memory>

#include "../../bpp_phyl_b200/host/bppgpu_shim.hpp"

using namespace bppshim;
using namespace std;

static void print_vec(const char* key, const Vdouble& v, bool comma = true) {
  printf("\"%s\": [", key);
  for (size_t i = 0; i < v.size(); ++i) printf("%s%.17g", i ? ", " : "", v[i]);
  printf("]%s\n", comma ? "," : "");
}
static void print_mat(const char* key, const RowMatrix<double>& m, bool comma = true) {
  printf("\"%s\": [", key);
  for (size_t i = 0; i < m.getNumberOfRows(); ++i) {
    printf("%s[", i ? ", " : "");
    for (size_t j = 0; j < m.getNumberOfColumns(); ++j) printf("%s%.17g", j ? ", " : "", m(i, j));
    printf("]");
  }
  printf("]%s\n", comma ? "," : "");
}
static void print_model(const char* name, const SubstitutionModel& m, bool comma = true) {
  printf("\"%s\": {\n", name);
  print_mat("Q", m.getGenerator());
  print_mat("V", m.getColumnRightEigenVectors());
  print_mat("Vinv", m.getRowLeftEigenVectors());
  print_vec("re", m.getEigenValues());
  print_vec("im", m.getIEigenValues());
  print_vec("freq", m.getFrequencies());
  printf("\"diagonalizable\": %d, \"nonsingular\": %d, \"rate\": %.17g\n}%s\n", (int)m.isDiagonalizable(), (int)m.isNonSingular(), m.getRate(),
         comma ? "," : "");
}

#include <fstream>
static double tree_len(const Node* n) {
  double s = n->hasFather() && n->hasDistanceToFather() ? n->getDistanceToFather() : 0.0;
  for (size_t i = 0; i < n->getNumberOfSons(); ++i) s += tree_len(n->getSon(i));
  return s;
}
static void io_checks() {
  // the input side: files -> tree / containers, the same objects the string constructors give
  const char* dir = getenv("BPPGPU_TEST_TMP") ? getenv("BPPGPU_TEST_TMP") : "/tmp";
  const std::string base = std::string(dir) + "/bppgpu_io_";
  { std::ofstream f(base + "t.nwk"); f << "((A:0.01,\n B[a comment]:0.02):0.03,\nC:0.01,D:0.1);\n(ignored);\n"; }
  { std::ofstream f(base + "s.fa"); f << ">A first\nAAATGG\nCTGTGC\n>B\naaatggctgtgg\n>C\nAAATGGCTGTGA\n>D\nNAATGG-TGTGC\n"; }
  { std::ofstream f(base + "s.phy"); f << " 2 6\nA  AAATGG\nB  CTG\nTGC\n"; }
  { std::ofstream f(base + "c.fa"); f << ">a\n7\n>b\n14\n\n>e\nX\n"; }
  unique_ptr<TreeTemplate<Node> > t(Newick(true).read(base + "t.nwk"));
  printf("\"IO_tree_leaves\": %zu, \"IO_tree_nodes\": %zu, \"IO_tree_len\": %.17g,\n", t->getLeavesNames().size(), t->getNodes().size(),
         tree_len(t->getRootNode()));
  VectorSiteContainer fa(&AlphabetTools::DNA_ALPHABET());
  Fasta().readSequences(base + "s.fa", fa);
  printf("\"IO_fasta_n\": %zu, \"IO_fasta_sites\": %zu, \"IO_fasta_name0\": \"%s\", \"IO_fasta_B7\": \"%s\",\n", fa.getNumberOfSequences(),
         fa.getNumberOfSites(), fa.getSequencesNames()[0].c_str(), fa.getSequence("B")[7].c_str());
  VectorSiteContainer ph(&AlphabetTools::DNA_ALPHABET());
  Phylip().readSequences(base + "s.phy", ph);
  printf("\"IO_phylip_n\": %zu, \"IO_phylip_B5\": \"%s\",\n", ph.getNumberOfSequences(), ph.getSequence("B")[5].c_str());
  ChromosomeAlphabet chr(1, 30);
  unique_ptr<VectorSiteContainer> cf(chrFasta::readSequencesFromFile(base + "c.fa", &chr));
  printf("\"IO_chr_n\": %zu, \"IO_chr_b\": \"%s\", \"IO_chr_e\": \"%s\",\n", cf->getNumberOfSequences(), cf->getSequence("b")[0].c_str(),
         cf->getSequence("e")[0].c_str());
  bool threw = false;
  try { Newick().read(base + "missing.nwk"); } catch (IOException&) { threw = true; }
  printf("\"IO_missing_throws\": %d,\n", (int)threw);
}

int main() {
  printf("{\n");
  io_checks();
  {
    GammaDiscreteRateDistribution g41(4, 1.0), g405(4, 0.5), g8(8, 2.3), g1(1, 0.7);
    Vdouble r;
    for (size_t i = 0; i < 4; ++i) r.push_back(g41.getCategory(i));
    print_vec("gamma_4_1", r);
    r.clear();
    for (size_t i = 0; i < 4; ++i) r.push_back(g405.getCategory(i));
    print_vec("gamma_4_0.5", r);
    r.clear();
    for (size_t i = 0; i < 8; ++i) r.push_back(g8.getCategory(i));
    print_vec("gamma_8_2.3", r);
    r.clear();
    r.push_back(g1.getCategory(0));
    print_vec("gamma_1_0.7", r);
  }
  const DNA* dna = &AlphabetTools::DNA_ALPHABET();
  {
    T92 t92(dna, 3.0, 0.5);
    print_model("T92", t92);
    GTR gtr(dna, 1.2, 0.8, 0.6, 1.5, 0.9, .3, .2, .25, .25);
    print_model("GTR", gtr);
    LG08 lg(&AlphabetTools::PROTEIN_ALPHABET());
    print_model("LG08", lg);
    YN98 yn(&AlphabetTools::CODON_ALPHABET(), 2.0, 0.3);
    print_model("YN98", yn);
    GY94 gy(&AlphabetTools::CODON_ALPHABET(), 2.0, 50.0);
    print_model("GY94", gy);
    ChromosomeAlphabet chr(1, 40);
    ChromosomeSubstitutionModel c1(&chr, 0.7, 0.4, 0.2, ChromosomeSubstitutionModel::DemiEqualDupl);
    print_model("CHR_REAL", c1);
    ChromosomeAlphabet chr2(1, 25);
    ChromosomeSubstitutionModel c2(&chr2, 1.5, 0.1, 0.9, 0.4, 0.05);
    print_model("CHR_COMPLEX", c2);
    ChromosomeAlphabet chr3(1, 20);
    ChromosomeSubstitutionModel c3(&chr3, 0.5, 0.0, 0.0, ChromosomeSubstitutionModel::IgnoreParam);
    print_model("CHR_SINGULAR", c3);
    // omega mixtures (YNGP_M2, RELAX): class probabilities, omegas and the synonymous-rate homogenisation, all host side
    {
      YNGP_M2 m2(&AlphabetTools::CODON_ALPHABET(), 2.0, 0.1, 2.0, 0.5, 0.8);
      RELAX rx(&AlphabetTools::CODON_ALPHABET(), 2.0, 0.1, 1.0, 2.0, 2.0, 0.5, 0.8);
      Vdouble r2, rr, q2;
      for (size_t k = 0; k < 3; ++k) {
        r2.push_back(m2.getNModel(k)->getRate());
        rr.push_back(rx.getNModel(k)->getRate());
        q2.push_back(m2.getNModel(k)->getGenerator()(2, 0));
      }
      print_vec("M2_rates", r2);
      print_vec("M2_probs", m2.getProbabilities());
      print_vec("M2_Q_AAG_AAA", q2);
      print_vec("RELAX_k2_rates", rr);
      SubstitutionModelSet* set = SubstitutionModelSetTools::createNonHomogeneousModelSet(
          new T92(&AlphabetTools::DNA_ALPHABET(), 3.), new GCFrequencySet(&AlphabetTools::DNA_ALPHABET()),
          unique_ptr<Tree>(TreeTemplateTools::parenthesisToTree("((A:0.1,B:0.2):0.3,(C:0.1,D:0.2):0.1);")).get(), {"T92.kappa"});
      set->setParameterValue("T92.kappa_1", 5.0);      // every copy follows through the aliases
      set->setParameterValue("T92.theta_4", 0.7);
      set->setParameterValue("GC.theta", 0.2);
      Vdouble kap, the;
      for (size_t k = 0; k < set->getNumberOfModels(); ++k) {
        kap.push_back(set->getModel(k)->getParameterValue("kappa"));
        the.push_back(set->getModel(k)->getParameterValue("theta"));
      }
      print_vec("NHSET_kappas", kap);
      print_vec("NHSET_thetas", the);
      print_vec("NHSET_rootfreqs", set->getRootFrequencies());
      printf("\"NHSET_nparams\": %zu,\n", set->getParameterNames().size());
      delete set;
    }
    // BatchedBrent (the lockstep line searches of the multi-start optimiser) on analytic functions: no device involved
    {
      struct Analytic {
        std::vector<double> x, y, val, a, b;
        unsigned calls = 0;
        size_t size() const { return x.size(); }
        double parameter(size_t k, const std::string& n) const { return n == "x" ? x[k] : y[k]; }
        void setParameter(size_t k, const std::string& n, double v) { (n == "x" ? x[k] : y[k]) = v; }
        void evaluate() {
          ++calls;
          for (size_t k = 0; k < x.size(); ++k)
            val[k] = (x[k] - a[k]) * (x[k] - a[k]) + 0.5 * std::pow(y[k] - b[k], 4) + 0.3 * std::sin(0.05 * x[k] * y[k]) + std::fabs(x[k] - 2 * a[k]);
        }
        double value(size_t k) const { return val[k]; }
      } fn;
      fn.a = {0.5, 3.0, 7.5, 20.0, 99.0};
      fn.b = {1.0, 0.2, 42.0, 5.5, 0.01};
      fn.x.assign(5, 10.0); fn.y.assign(5, 10.0); fn.val.assign(5, 0.0);
      fn.evaluate();
      Vdouble start = fn.val;
      BatchedBrent<Analytic> bb(&fn);
      std::vector<char> active = {1, 1, 1, 0, 1};        // point 3 is not searched: it must not move
      unsigned evals = 0;
      for (int round = 0; round < 12; ++round) {
        evals += bb.search("x", 1e-10, 100.0, 1e-8, active);
        evals += bb.search("y", 1e-10, 100.0, 1e-8, active);
      }
      print_vec("BRENT_x", fn.x);
      print_vec("BRENT_y", fn.y);
      print_vec("BRENT_val", fn.val);
      print_vec("BRENT_start", start);
      printf("\"BRENT_batch_evals\": %u, \"BRENT_calls\": %u,\n", evals, fn.calls);
    }
    // getInitValue / aliases
    printf("\"init_R\": [%g, %g, %g, %g],\n", t92.getInitValue(0, "R"), t92.getInitValue(1, "R"), t92.getInitValue(2, "R"), t92.getInitValue(3, "R"));
  }
  {
    // tree flattening: unrooting, post-order BrLen indexing, pre-order leaf order
    unique_ptr<Tree> t(TreeTemplateTools::parenthesisToTree("(((A:0.01, B:0.01):0.02,C:0.03):0.01,(D:0.04,E:0.05):0.06);"));
    t->unroot();
    t->resetNodesId();
    printf("\"unrooted_postorder\": [");
    bool first = true;
    for (Node* n : t->getNodes()) {
      printf("%s[\"%s\", %.17g, %d]", first ? "" : ", ", n->hasName() ? n->getName().c_str() : "", n->hasDistanceToFather() ? n->getDistanceToFather() : -1.0,
             n->hasFather() ? n->getFather()->getId() : -1);
      first = false;
    }
    printf("],\n\"leaves\": [");
    first = true;
    for (const string& s : t->getLeavesNames()) { printf("%s\"%s\"", first ? "" : ", ", s.c_str()); first = false; }
    printf("],\n");
  }
  {
    VectorSiteContainer sites(dna);
    sites.addSequence(BasicSequence("A", "AAATGGCTGTGCACGTC", dna));
    sites.addSequence(BasicSequence("B", "GACTGGATCTGCACGTC", dna));
    sites.addSequence(BasicSequence("C", "CTCTGGATGTGCACGTG", dna));
    sites.addSequence(BasicSequence("D", "AAATGGCGGTGCGCCTA", dna));
    SitePatterns sp(sites, {"A", "B", "C", "D"});
    printf("\"pattern_weights\": [");
    for (size_t i = 0; i < sp.getWeights().size(); ++i) printf("%s%u", i ? ", " : "", sp.getWeights()[i]);
    printf("],\n\"pattern_indices\": [");
    for (size_t i = 0; i < sp.getIndices().size(); ++i) printf("%s%ld", i ? ", " : "", (long)sp.getIndices()[i]);
    printf("],\n");
    // three-letter states: the codon alignment of test/test_relax.cpp with two columns repeated
    const CodonAlphabet* cod = &AlphabetTools::CODON_ALPHABET();
    VectorSiteContainer cs(cod);
    cs.addSequence(BasicSequence("A", "AAATGGCTGTGCACGTCTTGGAAA", cod));
    cs.addSequence(BasicSequence("B", "AACTGGATCTGCATGTCTTGGAAC", cod));
    cs.addSequence(BasicSequence("C", "ATCTGGACGTGCACGTGTTGGATC", cod));
    cs.addSequence(BasicSequence("D", "CAACGGGAGTGCGCCTATCGGCAA", cod));
    SitePatterns cp(cs, {"A", "B", "C", "D"});
    printf("\"codon_pattern_weights\": [");
    for (size_t i = 0; i < cp.getWeights().size(); ++i) printf("%s%u", i ? ", " : "", cp.getWeights()[i]);
    printf("],\n\"codon_pattern_indices\": [");
    for (size_t i = 0; i < cp.getIndices().size(); ++i) printf("%s%ld", i ? ", " : "", (long)cp.getIndices()[i]);
    printf("]\n");
  }
  printf("}\n");
  return 0;
}
