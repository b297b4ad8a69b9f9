// fasta.cpp -- rp_fasta_*: the reader of the many-pair front end (include/ractip_io.h).
// Behaviour follows Fasta::load of the reference (src/fa.cpp:37-83) line class by line class.
#include <cctype>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "ractip_io.h"
#include "ractip_prob.h"

struct rp_fasta {
  struct Rec { std::string name, seq, str; };
  std::vector<Rec> recs;
};

namespace {

const char kStructureChars[] = "()[].?xle ";

// an empty line counts as a structure line, as in the reference (its first "character" is the terminator, which
// strchr finds in any set)
bool structure_char(char ch) { return ch == '\0' || std::strchr(kStructureChars, ch) != nullptr; }

void parse_stream(std::istream& in, rp_fasta& f) {
  std::string line;
  rp_fasta::Rec cur;
  bool open = false;   // a header with a non-empty name has been seen
  while (std::getline(in, line)) {
    const char first = line.empty() ? '\0' : line[0];
    if (first == '>') {
      if (open) f.recs.push_back(cur);
      cur = rp_fasta::Rec();
      cur.name = line.substr(1);
      open = !cur.name.empty();
      continue;
    }
    size_t k = 0;
    if (!structure_char(first)) {
      while (k < line.size() && std::isalpha(static_cast<unsigned char>(line[k]))) k++;
      cur.seq.append(line, 0, k);
    } else {
      while (k < line.size() && line[k] != '\0' && std::strchr(kStructureChars, line[k]) != nullptr) k++;
      cur.str.append(line, 0, k);
    }
  }
  if (open) f.recs.push_back(cur);
}

}  // namespace

extern "C" {

int rp_fasta_load(const char* path, rp_fasta** out) {
  if (!path || !out) return RP_ERR_ARG;
  *out = nullptr;
  std::ifstream in(path);
  if (!in) return RP_ERR_ARG;
  rp_fasta* f = new rp_fasta();
  parse_stream(in, *f);
  *out = f;
  return RP_OK;
}

int rp_fasta_parse(const char* text, size_t len, rp_fasta** out) {
  if ((!text && len) || !out) return RP_ERR_ARG;
  std::istringstream in(std::string(text ? text : "", len));
  rp_fasta* f = new rp_fasta();
  parse_stream(in, *f);
  *out = f;
  return RP_OK;
}

int rp_fasta_count(const rp_fasta* f) { return f ? static_cast<int>(f->recs.size()) : 0; }

int rp_fasta_get(const rp_fasta* f, int k, const char** name, const char** seq, const char** str) {
  if (!f || k < 0 || k >= static_cast<int>(f->recs.size())) return RP_ERR_ARG;
  if (name) *name = f->recs[k].name.c_str();
  if (seq) *seq = f->recs[k].seq.c_str();
  if (str) *str = f->recs[k].str.c_str();
  return RP_OK;
}

void rp_fasta_free(rp_fasta* f) { delete f; }

}  // extern "C"
