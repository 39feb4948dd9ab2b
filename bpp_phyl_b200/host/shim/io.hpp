// bppgpu shim (see ../bppgpu_shim.hpp): the input side of the path: files -> TreeTemplate / VectorSiteContainer, the
// objects setData() turns into tip codes.  Newick (Io/Newick.cpp:69-92), and the sequence formats the reference's applications read:
// Fasta, sequential Phylip (bpp-seq) and the fork's chrFasta (one chromosome count per taxon, App/ChromosomeNumberMng.cpp:8).
#pragma once
#include <fstream>

#include "tree.hpp"

namespace bppshim {

class IOException : public Exception {
 public:
  explicit IOException(const std::string& m) : Exception(m) {}
};

// Io/Newick.cpp:69-92: lines are concatenated up to the first ';', bracketed comments removed, then parenthesisToTree
class Newick {
 public:
  explicit Newick(bool allowComments = false) : allowComments_(allowComments) {}
  TreeTemplate<Node>* readTree(std::istream& in) const {
    if (!in) throw IOException("Newick::read: failed to read from stream");
    std::string temp, description;
    while (std::getline(in, temp, '\n')) {
      const std::string::size_type index = temp.find(";");
      if (index != std::string::npos) { description += temp.substr(0, index + 1); break; }
      description += temp;
    }
    if (allowComments_) {
      std::string out;
      int depth = 0;
      for (char c : description) {
        if (c == '[') ++depth;
        else if (c == ']') { if (depth > 0) --depth; }
        else if (depth == 0) out += c;
      }
      description = out;
    }
    if (description.find_first_not_of(" \t\r\n") == std::string::npos) throw IOException("Newick::read: no tree was found!");
    return TreeTemplateTools::parenthesisToTree(description);
  }
  TreeTemplate<Node>* readTree(const std::string& path) const {
    std::ifstream in(path.c_str());
    if (!in) throw IOException("Newick::read: failed to read from stream");
    return readTree(in);
  }
  TreeTemplate<Node>* read(const std::string& path) const { return readTree(path); }

 private:
  bool allowComments_;
};

namespace iotools {
inline std::string strip(const std::string& s) {
  const size_t a = s.find_first_not_of(" \t\r\n");
  if (a == std::string::npos) return "";
  return s.substr(a, s.find_last_not_of(" \t\r\n") - a + 1);
}
inline std::string upper_nospace(const std::string& s) {
  std::string o;
  for (char c : s)
    if (!std::isspace((unsigned char)c)) o += (char)std::toupper((unsigned char)c);
  return o;
}
}  // namespace iotools

// bpp-seq Fasta: '>' name line (up to the first blank unless extended names), sequence over any number of lines
class Fasta {
 public:
  void readSequences(std::istream& in, VectorSiteContainer& sc) const {
    if (!in) throw IOException("Fasta::appendFromStream: can't read from istream input");
    std::string line, name, content;
    bool have = false;
    auto flush = [&]() {
      if (have) sc.addSequence(BasicSequence(name, iotools::upper_nospace(content), sc.getAlphabet()));
      content.clear();
    };
    while (std::getline(in, line)) {
      if (!line.empty() && line[0] == '>') {
        flush();
        name = iotools::strip(line.substr(1));
        have = true;
      } else if (have) {
        content += line;
      }
    }
    flush();
  }
  void readSequences(const std::string& path, VectorSiteContainer& sc) const {
    std::ifstream in(path.c_str());
    if (!in) throw IOException("Fasta::readSequences: failed to open " + path);
    readSequences(in, sc);
  }
};

// bpp-seq Phylip, sequential: "<n> <len>" then n records "name  sequence" (sequence may continue on following lines)
class Phylip {
 public:
  void readSequences(std::istream& in, VectorSiteContainer& sc) const {
    size_t n = 0, len = 0;
    if (!(in >> n >> len)) throw IOException("Phylip::read: bad header");
    const unsigned w = std::max(1u, sc.getAlphabet()->getStateCodingSize());
    for (size_t k = 0; k < n; ++k) {
      std::string name, content, tok;
      if (!(in >> name)) throw IOException("Phylip::read: missing sequence");
      while (content.size() < len * w && (in >> tok)) content += tok;
      if (content.size() != len * w) throw IOException("Phylip::read: sequence " + name + " has a wrong length");
      sc.addSequence(BasicSequence(name, iotools::upper_nospace(content), sc.getAlphabet()));
    }
  }
  void readSequences(const std::string& path, VectorSiteContainer& sc) const {
    std::ifstream in(path.c_str());
    if (!in) throw IOException("Phylip::readSequences: failed to open " + path);
    readSequences(in, sc);
  }
};

// the fork's chrFasta (bpp-seq, not under /root/reference; used as chrFasta::readSequencesFromFile(path, alphabet) in
// App/ChromosomeNumberMng.cpp:8): Fasta whose "sequence" is ONE chromosome count per taxon ("X" = unknown)
class chrFasta {
 public:
  static VectorSiteContainer* readSequencesFromFile(const std::string& path, const ChromosomeAlphabet* alpha) {
    std::ifstream in(path.c_str());
    if (!in) throw IOException("chrFasta::readSequencesFromFile: failed to open " + path);
    VectorSiteContainer* sc = new VectorSiteContainer(alpha);
    std::string line, name;
    while (std::getline(in, line)) {
      line = iotools::strip(line);
      if (line.empty()) continue;
      if (line[0] == '>') name = iotools::strip(line.substr(1));
      else if (!name.empty()) { sc->addSequence(BasicSequence(name, line, alpha)); name.clear(); }
    }
    return sc;
  }
};

}  // namespace bppshim
