// model.cpp -- host side of the energy-model layer: where the integer tables
// (rp_model) come from.
//
// Replaces, for the probability stage:
//   * copy_boltzmann_parameters()   reference src/boltzmann_param.c:5908-6026
//     (Andronescu BL* values written over ViennaRNA's Turner-2004 globals)
//   * Vienna::read_parameter_file() reference src/ractip.cpp:63,1568-1569
//     (ViennaRNA "RNAfold parameter file v2.0" reader)
// in the order RactIP::run applies them (src/ractip.cpp:1566-1569): defaults,
// then BL*, then the -P file.
//
// The BL* numbers are embedded from params/blstar.bin and the Turner-2004
// residual tables from params/turner2004_residual.par (see that file for
// provenance); both are turned into a C array by ractip_b200/build.py.
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ractip_prob.h"

extern "C" {
extern const unsigned char rp_embedded_blstar[];
extern const unsigned int rp_embedded_blstar_len;
extern const char rp_embedded_residual_par[];
}

namespace {

constexpr int NB = RP_NBPAIRS;

// ---------------------------------------------------------------- BL* blob
struct BlArray { std::string name; std::vector<int> v; };

bool parse_blob(std::vector<BlArray>& arrays, std::vector<std::pair<std::string, int>>& tetra) {
  const unsigned char* p = rp_embedded_blstar;
  const unsigned char* end = p + rp_embedded_blstar_len;
  if (rp_embedded_blstar_len < 12 || std::memcmp(p, "RPBLSTR1", 8) != 0) return false;
  p += 8;
  auto rd32 = [&](int& out) {
    if (p + 4 > end) return false;
    int32_t x;
    std::memcpy(&x, p, 4);
    p += 4;
    out = x;
    return true;
  };
  int na;
  if (!rd32(na)) return false;
  for (int a = 0; a < na; a++) {
    if (p + 24 > end) return false;
    BlArray arr;
    arr.name = std::string(reinterpret_cast<const char*>(p));
    p += 24;
    int cnt;
    if (!rd32(cnt)) return false;
    arr.v.resize(cnt);
    for (int i = 0; i < cnt; i++)
      if (!rd32(arr.v[i])) return false;
    arrays.push_back(std::move(arr));
  }
  int nt;
  if (!rd32(nt)) return false;
  for (int t = 0; t < nt; t++) {
    if (p + 12 > end) return false;
    std::string s(reinterpret_cast<const char*>(p));
    p += 8;
    int e;
    rd32(e);
    tetra.emplace_back(s, e);
  }
  return true;
}

const std::vector<int>* find(const std::vector<BlArray>& arrays, const char* name, size_t n) {
  for (auto& a : arrays)
    if (a.name == name && a.v.size() == n) return &a.v;
  return nullptr;
}

// The index ranges below are the reference's copy semantics, which matter:
// copy_stacks/int11/int21 fill pair types 1..7, copy_int22 fills bases 1..4
// only, copy_dangle fills rows 0..7, copy_mismatch rows 1..7
// (src/boltzmann_param.c:5908-5971).
int apply_blstar(rp_model* m) {
  std::vector<BlArray> arrays;
  std::vector<std::pair<std::string, int>> tetra;
  if (!parse_blob(arrays, tetra)) return RP_ERR_FORMAT;
  const std::vector<int>*stack = find(arrays, "stack37a", 49), *mmH = find(arrays, "mismatchH37a", 175),
                         *mmI = find(arrays, "mismatchI37a", 175), *d5 = find(arrays, "dangle5_37a", 40),
                         *d3 = find(arrays, "dangle3_37a", 40), *i11 = find(arrays, "int11_37a", 1225),
                         *i21 = find(arrays, "int21_37a", 6125), *i22 = find(arrays, "int22_37a", 12544),
                         *hp = find(arrays, "hairpin37a", 31), *bu = find(arrays, "bulge37a", 31),
                         *il = find(arrays, "internal_loop37a", 31), *ml = find(arrays, "MLparams_a", 4),
                         *ni = find(arrays, "ninio_a", 2);
  if (!stack || !mmH || !mmI || !d5 || !d3 || !i11 || !i21 || !i22 || !hp || !bu || !il || !ml || !ni)
    return RP_ERR_FORMAT;
  int p = 0;
  for (int i = 1; i <= NB; i++)
    for (int j = 1; j <= NB; j++) m->stack37[i][j] = (*stack)[p++];
  p = 0;
  for (int i = 1; i <= NB; i++)
    for (int j = 0; j < 5; j++)
      for (int k = 0; k < 5; k++, p++) {
        m->mismatchH37[i][j][k] = (*mmH)[p];
        m->mismatchI37[i][j][k] = (*mmI)[p];
      }
  p = 0;
  for (int i = 0; i <= NB; i++)
    for (int j = 0; j < 5; j++, p++) {
      m->dangle5_37[i][j] = (*d5)[p];
      m->dangle3_37[i][j] = (*d3)[p];
    }
  p = 0;
  for (int i = 1; i <= NB; i++)
    for (int j = 1; j <= NB; j++)
      for (int k = 0; k < 5; k++)
        for (int l = 0; l < 5; l++) m->int11_37[i][j][k][l] = (*i11)[p++];
  p = 0;
  for (int i = 1; i <= NB; i++)
    for (int j = 1; j <= NB; j++)
      for (int k = 0; k < 5; k++)
        for (int l = 0; l < 5; l++)
          for (int a = 0; a < 5; a++) m->int21_37[i][j][k][l][a] = (*i21)[p++];
  p = 0;
  for (int i = 1; i <= NB; i++)
    for (int j = 1; j <= NB; j++)
      for (int k = 1; k < 5; k++)
        for (int l = 1; l < 5; l++)
          for (int a = 1; a < 5; a++)
            for (int b = 1; b < 5; b++) m->int22_37[i][j][k][l][a][b] = (*i22)[p++];
  for (int i = 0; i < 31; i++) {
    m->hairpin37[i] = (*hp)[i];
    m->bulge37[i] = (*bu)[i];
    m->internal_loop37[i] = (*il)[i];
  }
  m->ML_BASE37 = (*ml)[0];
  m->ML_closing37 = (*ml)[1];
  m->ML_intern37 = (*ml)[2];
  m->TerminalAU37 = (*ml)[3];
  m->ninio37 = (*ni)[0];
  m->MAX_NINIO = (*ni)[1];
  // copy_Tetra_loop: strcpy at 7*i then strcat(" ") truncates the list to the
  // entries written so far, so the default list is fully replaced
  // (src/boltzmann_param.c:5995-6008).
  for (size_t i = 0; i < tetra.size() && i < 200; i++) {
    std::strcpy(&m->Tetraloops[7 * i], tetra[i].first.c_str());
    std::strcat(m->Tetraloops, " ");
    m->Tetraloop37[i] = tetra[i].second;
  }
  return RP_OK;
}

// ------------------------------------------------------------- .par reader
// Strip /* */ comments (may span lines) and split the file into sections.
struct Section { std::string name; std::vector<std::string> lines; };

int split_sections(const std::string& text, std::vector<Section>& out) {
  std::string clean;
  clean.reserve(text.size());
  for (size_t i = 0; i < text.size();) {
    if (text.compare(i, 2, "/*") == 0) {
      size_t e = text.find("*/", i + 2);
      if (e == std::string::npos) return RP_ERR_FORMAT;
      i = e + 2;
      clean.push_back(' ');
    } else {
      clean.push_back(text[i++]);
    }
  }
  size_t i = 0;
  bool header_ok = false;
  Section* cur = nullptr;
  while (i < clean.size()) {
    size_t e = clean.find('\n', i);
    if (e == std::string::npos) e = clean.size();
    std::string line = clean.substr(i, e - i);
    i = e + 1;
    size_t a = line.find_first_not_of(" \t\r");
    if (a == std::string::npos) continue;
    if (line.compare(a, 2, "##") == 0) {
      if (line.find("RNAfold parameter file v2.0") != std::string::npos) header_ok = true;
      continue;
    }
    if (line[a] == '#') {
      size_t b = line.find_first_not_of(" \t", a + 1);
      std::string name = b == std::string::npos ? "" : line.substr(b);
      while (!name.empty() && std::isspace((unsigned char)name.back())) name.pop_back();
      out.push_back(Section{name, {}});
      cur = &out.back();
      continue;
    }
    if (cur) cur->lines.push_back(line);
  }
  return header_ok ? RP_OK : RP_ERR_FORMAT;
}

bool parse_ints(const Section& s, std::vector<int>& v) {
  for (auto& line : s.lines) {
    size_t i = 0;
    while (i < line.size()) {
      while (i < line.size() && std::isspace((unsigned char)line[i])) i++;
      if (i >= line.size()) break;
      size_t e = i;
      while (e < line.size() && !std::isspace((unsigned char)line[e])) e++;
      std::string tok = line.substr(i, e - i);
      i = e;
      if (tok == "INF") v.push_back(RP_INF);
      else if (tok == "DEF") v.push_back(-50);
      else if (tok == "NST") v.push_back(0);
      else {
        char* endp = nullptr;
        long x = std::strtol(tok.c_str(), &endp, 10);
        if (*endp != 0) {
          // Misc carries one double (lxc); the caller handles that section itself
          return false;
        }
        v.push_back((int)x);
      }
    }
  }
  return true;
}

int rd_mismatch(const Section& s, int arr[NB + 1][5][5]) {
  std::vector<int> v;
  if (!parse_ints(s, v) || v.size() != (size_t)NB * 25) return RP_ERR_FORMAT;
  size_t p = 0;
  for (int i = 1; i <= NB; i++)
    for (int j = 0; j < 5; j++)
      for (int k = 0; k < 5; k++) arr[i][j][k] = v[p++];
  return RP_OK;
}

int rd_dangle(const Section& s, int arr[NB + 1][5]) {
  std::vector<int> v;
  if (!parse_ints(s, v) || v.size() != (size_t)NB * 5) return RP_ERR_FORMAT;
  size_t p = 0;
  for (int i = 1; i <= NB; i++)
    for (int j = 0; j < 5; j++) arr[i][j] = v[p++];
  return RP_OK;
}

int rd_loop31(const Section& s, int arr[31]) {
  std::vector<int> v;
  if (!parse_ints(s, v) || v.size() != 31) return RP_ERR_FORMAT;
  for (int i = 0; i < 31; i++) arr[i] = v[i];
  return RP_OK;
}

// Entries of int22 that involve an unknown base (index 0) are not in the file;
// fill them with the least favourable value over ACGU in that slot.
void int22_fill_unknown(rp_model* m) {
  for (int i = 1; i <= NB; i++)
    for (int j = 1; j <= NB; j++)
      for (int mask = 1; mask < 16; mask++)  // which of the 4 base slots are N
        for (int k = 0; k < 5; k++)
          for (int l = 0; l < 5; l++)
            for (int a = 0; a < 5; a++)
              for (int b = 0; b < 5; b++) {
                int idx[4] = {k, l, a, b};
                bool match = true;
                for (int s = 0; s < 4; s++)
                  if (((mask >> s) & 1) != (idx[s] == 0)) match = false;
                if (!match) continue;
                int best = -RP_INF;
                int lo[4], hi[4];
                for (int s = 0; s < 4; s++) {
                  lo[s] = idx[s] == 0 ? 1 : idx[s];
                  hi[s] = idx[s] == 0 ? 4 : idx[s];
                }
                for (int x0 = lo[0]; x0 <= hi[0]; x0++)
                  for (int x1 = lo[1]; x1 <= hi[1]; x1++)
                    for (int x2 = lo[2]; x2 <= hi[2]; x2++)
                      for (int x3 = lo[3]; x3 <= hi[3]; x3++)
                        if (m->int22_37[i][j][x0][x1][x2][x3] > best) best = m->int22_37[i][j][x0][x1][x2][x3];
                m->int22_37[i][j][k][l][a][b] = best;
              }
}

int rd_special(const Section& s, char* names, size_t names_cap, int* en, int max_entries, int mer) {
  std::memset(names, 0, names_cap);
  std::memset(en, 0, sizeof(int) * max_entries);
  int n = 0;
  for (auto& line : s.lines) {
    char buf[32];
    int e37, eH;
    if (std::sscanf(line.c_str(), "%31s %d %d", buf, &e37, &eH) < 2) continue;
    if ((int)std::strlen(buf) != mer || n >= max_entries) return RP_ERR_FORMAT;
    if ((size_t)((n + 1) * (mer + 1)) >= names_cap) return RP_ERR_FORMAT;
    std::strcat(names, buf);
    std::strcat(names, " ");
    en[n++] = e37;
  }
  return RP_OK;
}

int apply_par_text(rp_model* m, const std::string& text) {
  std::vector<Section> secs;
  int rc = split_sections(text, secs);
  if (rc) return rc;
  for (auto& s : secs) {
    const std::string& n = s.name;
    std::vector<int> v;
    if (n == "END") break;
    // enthalpy tables cancel at 37 C and are not part of rp_model
    if (n.size() > 11 && n.compare(n.size() - 11, 11, "_enthalpies") == 0) continue;
    if (n == "stack") {
      if (!parse_ints(s, v) || v.size() != 49) return RP_ERR_FORMAT;
      size_t p = 0;
      for (int i = 1; i <= NB; i++)
        for (int j = 1; j <= NB; j++) m->stack37[i][j] = v[p++];
    } else if (n == "mismatch_hairpin") rc = rd_mismatch(s, m->mismatchH37);
    else if (n == "mismatch_interior") rc = rd_mismatch(s, m->mismatchI37);
    else if (n == "mismatch_interior_1n") rc = rd_mismatch(s, m->mismatch1nI37);
    else if (n == "mismatch_interior_23") rc = rd_mismatch(s, m->mismatch23I37);
    else if (n == "mismatch_multi") rc = rd_mismatch(s, m->mismatchM37);
    else if (n == "mismatch_exterior") rc = rd_mismatch(s, m->mismatchExt37);
    else if (n == "dangle5") rc = rd_dangle(s, m->dangle5_37);
    else if (n == "dangle3") rc = rd_dangle(s, m->dangle3_37);
    else if (n == "int11") {
      if (!parse_ints(s, v) || v.size() != 1225) return RP_ERR_FORMAT;
      size_t p = 0;
      for (int i = 1; i <= NB; i++)
        for (int j = 1; j <= NB; j++)
          for (int k = 0; k < 5; k++)
            for (int l = 0; l < 5; l++) m->int11_37[i][j][k][l] = v[p++];
    } else if (n == "int21") {
      if (!parse_ints(s, v) || v.size() != 6125) return RP_ERR_FORMAT;
      size_t p = 0;
      for (int i = 1; i <= NB; i++)
        for (int j = 1; j <= NB; j++)
          for (int k = 0; k < 5; k++)
            for (int l = 0; l < 5; l++)
              for (int a = 0; a < 5; a++) m->int21_37[i][j][k][l][a] = v[p++];
    } else if (n == "int22") {
      // v2.0 files carry the 6x6 canonical pair types x 4^4 known bases
      if (!parse_ints(s, v) || v.size() != 36 * 256) return RP_ERR_FORMAT;
      size_t p = 0;
      for (int i = 1; i < NB; i++)
        for (int j = 1; j < NB; j++)
          for (int k = 1; k < 5; k++)
            for (int l = 1; l < 5; l++)
              for (int a = 1; a < 5; a++)
                for (int b = 1; b < 5; b++) m->int22_37[i][j][k][l][a][b] = v[p++];
      int22_fill_unknown(m);
    } else if (n == "hairpin") rc = rd_loop31(s, m->hairpin37);
    else if (n == "bulge") rc = rd_loop31(s, m->bulge37);
    else if (n == "interior") rc = rd_loop31(s, m->internal_loop37);
    else if (n == "NINIO") {
      if (!parse_ints(s, v) || v.size() < 3) return RP_ERR_FORMAT;
      m->ninio37 = v[0];
      m->MAX_NINIO = v[2];
    } else if (n == "ML_params") {
      if (!parse_ints(s, v) || v.size() < 6) return RP_ERR_FORMAT;
      m->ML_BASE37 = v[0];
      m->ML_closing37 = v[2];
      m->ML_intern37 = v[4];
    } else if (n == "Misc") {
      int di = 0, dih = 0, tau = 0, tauh = 0;
      double lxc = 0;
      bool ok = false;
      for (auto& line : s.lines) {
        int got = std::sscanf(line.c_str(), "%d %d %d %d %lf", &di, &dih, &tau, &tauh, &lxc);
        if (got >= 4) {
          m->DuplexInit37 = di;
          m->TerminalAU37 = tau;
          if (got >= 5) m->lxc37 = lxc;
          ok = true;
          break;
        }
      }
      if (!ok) return RP_ERR_FORMAT;
    } else if (n == "Tetraloops") rc = rd_special(s, m->Tetraloops, sizeof m->Tetraloops, m->Tetraloop37, 200, 6);
    else if (n == "Triloops") rc = rd_special(s, m->Triloops, sizeof m->Triloops, m->Triloop37, 40, 5);
    else if (n == "Hexaloops") rc = rd_special(s, m->Hexaloops, sizeof m->Hexaloops, m->Hexaloop37, 200, 8);
    // unknown sections are ignored, as ViennaRNA does (it warns)
    if (rc) return rc;
  }
  return RP_OK;
}

}  // namespace

extern "C" {

int rp_model_default(rp_model* m, int use_bl) {
  if (!m) return RP_ERR_ARG;
  std::memset(m, 0, sizeof *m);
  m->temperature = 37.0;
  m->dangles = 2;
  m->special_hp = 1;
  m->pf_smooth = 1;
  m->sfact = 1.07;
  int rc = apply_par_text(m, rp_embedded_residual_par);
  if (rc) return rc;
  if (!use_bl) return RP_ERR_NO_DEFAULTS;  // full Turner-2004 set is not embedded
  return apply_blstar(m);
}

int rp_model_read_par(rp_model* m, const char* path) {
  if (!m || !path) return RP_ERR_ARG;
  FILE* f = std::fopen(path, "rb");
  if (!f) return RP_ERR_IO;
  std::string text;
  char buf[65536];
  size_t n;
  while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
  std::fclose(f);
  return apply_par_text(m, text);
}

uint64_t rp_model_digest(const rp_model* m) {
  if (!m) return 0;
  // hash field by field so struct padding never leaks in
  uint64_t h = 1469598103934665603ull;
  auto mix = [&](const void* p, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; i++) {
      h ^= b[i];
      h *= 1099511628211ull;
    }
  };
#define MIX(f) mix(&m->f, sizeof m->f)
  MIX(temperature); MIX(dangles); MIX(special_hp); MIX(pf_smooth); MIX(sfact);
  MIX(stack37); MIX(hairpin37); MIX(bulge37); MIX(internal_loop37);
  MIX(mismatchI37); MIX(mismatchH37); MIX(mismatchM37); MIX(mismatchExt37);
  MIX(mismatch1nI37); MIX(mismatch23I37); MIX(dangle5_37); MIX(dangle3_37);
  MIX(int11_37); MIX(int21_37); MIX(int22_37);
  MIX(ML_BASE37); MIX(ML_closing37); MIX(ML_intern37); MIX(TerminalAU37);
  MIX(ninio37); MIX(MAX_NINIO); MIX(DuplexInit37); MIX(lxc37);
  MIX(Tetraloops); MIX(Tetraloop37); MIX(Triloops); MIX(Triloop37);
  MIX(Hexaloops); MIX(Hexaloop37);
#undef MIX
  return h;
}

}  // extern "C"
