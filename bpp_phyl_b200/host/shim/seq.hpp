// bppgpu shim (see ../bppgpu_shim.hpp): alphabets, sequences, site containers (bpp-seq stand-ins)
#pragma once
#include "core.hpp"

namespace bppshim {

// ---- bpp-seq stand-ins ------------------------------------------------------------------------------------------
class Alphabet {
 public:
  virtual ~Alphabet() {}
  virtual size_t getSize() const = 0;                         // number of resolved states
  virtual unsigned getStateCodingSize() const { return 1; }   // characters per state
  // resolved states a character (string of getStateCodingSize() chars) stands for; empty = unknown character
  virtual std::vector<int> getAlias(const std::string& ch) const = 0;
};
class LetterAlphabet : public Alphabet {
 public:
  LetterAlphabet(const std::string& states, const std::map<char, std::string>& aliases) : states_(states), aliases_(aliases) {}
  size_t getSize() const { return states_.size(); }
  std::vector<int> getAlias(const std::string& ch) const {
    std::vector<int> out;
    if (ch.size() != 1) return out;
    const char c = (char)std::toupper((unsigned char)ch[0]);
    const size_t p = states_.find(c);
    if (p != std::string::npos) { out.push_back((int)p); return out; }
    std::map<char, std::string>::const_iterator it = aliases_.find(c);
    if (it != aliases_.end())
      for (char r : it->second) out.push_back((int)states_.find(r));
    return out;
  }
  const std::string& states() const { return states_; }

 private:
  std::string states_;
  std::map<char, std::string> aliases_;
};
class DNA : public LetterAlphabet {
 public:
  DNA() : LetterAlphabet("ACGT", {{'U', "T"}, {'M', "AC"}, {'R', "AG"}, {'W', "AT"}, {'S', "CG"}, {'Y', "CT"}, {'K', "GT"},
                                  {'V', "ACG"}, {'H', "ACT"}, {'D', "AGT"}, {'B', "CGT"}, {'N', "ACGT"}, {'X', "ACGT"},
                                  {'O', "ACGT"}, {'0', "ACGT"}, {'?', "ACGT"}, {'-', "ACGT"}}) {}
};
class ProteicAlphabet : public LetterAlphabet {
 public:
  ProteicAlphabet() : LetterAlphabet("ARNDCQEGHILKMFPSTWYV", {{'B', "ND"}, {'Z', "QE"}, {'J', "IL"},
                                                               {'X', "ARNDCQEGHILKMFPSTWYV"}, {'O', "ARNDCQEGHILKMFPSTWYV"},
                                                               {'0', "ARNDCQEGHILKMFPSTWYV"}, {'?', "ARNDCQEGHILKMFPSTWYV"},
                                                               {'-', "ARNDCQEGHILKMFPSTWYV"}}) {}
};
// 64 codons, index 16*n1 + 4*n2 + n3 with A,C,G,T = 0..3; standard genetic code
class CodonAlphabet : public Alphabet {
 public:
  size_t getSize() const { return 64; }
  unsigned getStateCodingSize() const { return 3; }
  std::vector<int> getAlias(const std::string& ch) const {
    std::vector<int> out;
    if (ch.size() != 3) return out;
    std::vector<int> pos[3];
    DNA dna;
    for (int k = 0; k < 3; ++k) pos[k] = dna.getAlias(std::string(1, ch[k]));
    for (int a : pos[0]) for (int b : pos[1]) for (int c : pos[2]) out.push_back(16 * a + 4 * b + c);
    return out;
  }
  static char aminoAcid(int codon) {
    static const char* tcag = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    static const int map_[4] = {2, 1, 3, 0};  // A,C,G,T -> position in T,C,A,G
    const int a = codon / 16, b = (codon / 4) % 4, c = codon % 4;
    return tcag[16 * map_[a] + 4 * map_[b] + map_[c]];
  }
  static bool isStop(int codon) { return aminoAcid(codon) == '*'; }
};
// chromosome counts min..max, one integer per taxon written in decimal ("X" = unknown)
class ChromosomeAlphabet : public Alphabet {
 public:
  ChromosomeAlphabet(unsigned mn, unsigned mx) : min_(mn), max_(mx) {}
  size_t getSize() const { return max_ - min_ + 1; }
  unsigned getStateCodingSize() const { return 0; }  // variable width: whitespace-separated counts
  unsigned getMin() const { return min_; }
  unsigned getMax() const { return max_; }
  std::vector<int> getAlias(const std::string& ch) const {
    std::vector<int> out;
    if (ch == "X" || ch == "x" || ch == "-" || ch == "?") {
      for (unsigned s = 0; s < getSize(); ++s) out.push_back((int)s);
      return out;
    }
    const long v = std::strtol(ch.c_str(), nullptr, 10);
    if (v >= (long)min_ && v <= (long)max_) out.push_back((int)(v - min_));
    return out;
  }

 private:
  unsigned min_, max_;
};
namespace AlphabetTools {
inline const DNA& DNA_ALPHABET() { static DNA a; return a; }
inline const ProteicAlphabet& PROTEIN_ALPHABET() { static ProteicAlphabet a; return a; }
inline const CodonAlphabet& CODON_ALPHABET() { static CodonAlphabet a; return a; }
}  // namespace AlphabetTools

class BasicSequence {
 public:
  BasicSequence(const std::string& name, const std::string& content, const Alphabet* alpha) : name_(name), alpha_(alpha) {
    const unsigned w = alpha->getStateCodingSize();
    if (w == 0) {  // variable width: whitespace separated (chromosome counts)
      std::istringstream is(content);
      std::string tok;
      while (is >> tok) states_.push_back(tok);
    } else {
      for (size_t i = 0; i + w <= content.size(); i += w) states_.push_back(content.substr(i, w));
    }
  }
  BasicSequence(const std::string& name, const std::vector<std::string>& states, const Alphabet* alpha) : name_(name), states_(states), alpha_(alpha) {}
  const std::string& getName() const { return name_; }
  size_t size() const { return states_.size(); }
  const std::string& operator[](size_t i) const { return states_[i]; }
  const Alphabet* getAlphabet() const { return alpha_; }

 private:
  std::string name_;
  std::vector<std::string> states_;
  const Alphabet* alpha_;
};

class VectorSiteContainer {
 public:
  explicit VectorSiteContainer(const Alphabet* alpha) : alpha_(alpha) {}
  void addSequence(const BasicSequence& s) {
    if (!seqs_.empty() && s.size() != seqs_[0].size()) throw Exception("VectorSiteContainer::addSequence. Sequence " + s.getName() + " has a different length.");
    seqs_.push_back(s);
  }
  size_t getNumberOfSequences() const { return seqs_.size(); }
  size_t getNumberOfSites() const { return seqs_.empty() ? 0 : seqs_[0].size(); }
  const Alphabet* getAlphabet() const { return alpha_; }
  std::vector<std::string> getSequencesNames() const {
    std::vector<std::string> n;
    for (const BasicSequence& s : seqs_) n.push_back(s.getName());
    return n;
  }
  const BasicSequence& getSequence(const std::string& name) const {
    for (const BasicSequence& s : seqs_)
      if (s.getName() == name) return s;
    throw Exception("SequenceNotFoundException: " + name);
  }

 private:
  const Alphabet* alpha_;
  std::vector<BasicSequence> seqs_;
};

}  // namespace bppshim
