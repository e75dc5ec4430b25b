// chem_loader.cpp -- C++ host-side mirror of the reference's Fortran loaders for the
// chemistry path (the Fortran toolchain is absent in this image, so the host side
// above the C-ABI is written in C++ with the reference's subroutine names):
//   chem_read_reactions / chem_load_reactions   src/chemistry.f90:1427-1454, 1364-1424
//   chem_parse_reactions / getElements          src/chemistry.f90:1221-1360, 1458-1529
//   chem_get_dupli_reactions                    src/chemistry.f90:1188-1217
//   chem_load_initial_abundances                src/chemistry.f90:1978-2024
// In production these stay Fortran (north_star); this mirror feeds the same tables
// to racg_network_create for the benchmarks, the tests and the C harness.
// Written independently of oracle/ (hash-map species discovery, tokenising element
// parser) so that comparing the two is a meaningful parity test.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../include/racg.h"

namespace {

struct ChemHost {
  int R = 0, N = 0;
  std::vector<int> reac, prod, n_reac, n_prod, itype;
  std::vector<double> ABC, T_range;
  std::vector<char> ctype;            // 2*R
  std::vector<std::string> names;
  std::vector<int> elements;          // 20*N
  std::vector<double> mass_num, vib_freq, Edesorb;
  std::vector<int> dupli_ptr, dupli_list;
  std::string err;
};

const char* ELEM[RACG_NELEM] = {"+-", "E", "Grain", "H", "D", "He", "C", "N", "O", "Si", "S", "Fe",
                                "Na", "Mg", "Cl", "P", "F", "Ne", "Ar", "K"};
const double EMASS[RACG_NELEM] = {0.0, 5.45e-4, 0.0, 1.0, 2.0, 4.0, 12.0, 14.0, 16.0, 28.0, 32.0, 56.0,
                                  23.0, 24.0, 35.5, 31.0, 19.0, 20.18, 39.95, 39.1};

std::string strip(const std::string& s) {
  size_t a = s.find_first_not_of(' ');
  if (a == std::string::npos) return "";
  size_t b = s.find_last_not_of(' ');
  return s.substr(a, b - a + 1);
}

// value of a fixed-width Fortran Fw.0 field
double field_real(const std::string& raw) {
  std::string t;
  for (char c : raw) if (c != ' ') t.push_back((c == 'D' || c == 'd') ? 'E' : c);
  if (t.empty()) return 0.0;
  // "1.5-3" form: a sign that is not first and does not follow an exponent letter
  for (size_t i = 1; i < t.size(); ++i)
    if ((t[i] == '+' || t[i] == '-') && t[i - 1] != 'E' && t[i - 1] != 'e') { t.insert(i, "E"); break; }
  return strtod(t.c_str(), nullptr);
}

// elemental composition: longest-match tokenizer (Grain, two-letter symbols, then
// one-letter symbols), a following integer multiplies the preceding element, a
// trailing +/- sets the charge slot
void compose(const std::string& nm, int* e) {
  for (int k = 0; k < RACG_NELEM; ++k) e[k] = 0;
  int last = -1;
  size_t i = 0;
  while (i < nm.size()) {
    int hit = -1; size_t hl = 0;
    for (int k = 1; k < RACG_NELEM; ++k) {
      size_t l = strlen(ELEM[k]);
      if (l > hl && nm.compare(i, l, ELEM[k]) == 0) { hit = k; hl = l; }
    }
    if (hit >= 0) { e[hit] += 1; last = hit; i += hl; continue; }
    char c = nm[i];
    if (isdigit((unsigned char)c) && i > 0) {
      size_t j = i; int v = 0;
      while (j < nm.size() && j < i + 2 && isdigit((unsigned char)nm[j])) { v = v * 10 + (nm[j] - '0'); ++j; }
      if (v != 0 && last >= 0) e[last] += v - 1;
      // further digits (3+) are ignored, as the reference reads at most two
      while (j < nm.size() && isdigit((unsigned char)nm[j])) ++j;
      i = j; continue;
    }
    if (c == '+' && i > 0) e[0] = 1;
    else if (c == '-' && i > 0) e[0] = -1;
    ++i;
  }
}

bool read_network(ChemHost& h, const char* file) {
  std::ifstream in(file);
  if (!in) { h.err = std::string("cannot open ") + file; return false; }
  std::string line;
  std::vector<std::string> rows;
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty() || line[0] == '!' || line[0] == ' ') continue;
    if (line.size() < 150) line.append(150 - line.size(), ' ');
    rows.push_back(line);
  }
  const int R = h.R = (int)rows.size();
  h.reac.assign(3 * R, 0); h.prod.assign(4 * R, 0); h.n_reac.assign(R, 0); h.n_prod.assign(R, 0);
  h.itype.assign(R, 0); h.ABC.assign(3 * R, 0); h.T_range.assign(2 * R, 0); h.ctype.assign(2 * R, ' ');
  std::unordered_map<std::string, int> id;
  h.names.clear();
  auto species = [&](const std::string& s) {
    auto it = id.find(s);
    if (it != id.end()) return it->second;
    h.names.push_back(s);
    id[s] = (int)h.names.size();
    return (int)h.names.size();
  };
  for (int i = 0; i < R; ++i) {
    const std::string& s = rows[i];
    std::string rn[3], pn[4];
    for (int k = 0; k < 3; ++k) rn[k] = strip(s.substr(12 * k, 12));
    for (int k = 0; k < 4; ++k) pn[k] = strip(s.substr(36 + 12 * k, 12));
    for (int k = 0; k < 3; ++k) h.ABC[3 * i + k] = field_real(s.substr(84 + 9 * k, 9));
    for (int k = 0; k < 2; ++k) h.T_range[2 * i + k] = field_real(s.substr(111 + 6 * k, 6));
    h.itype[i] = (int)field_real(s.substr(123, 3));
    h.ctype[2 * i] = s[129]; h.ctype[2 * i + 1] = s[130];
    int nr = 0, np = 0;
    for (int k = 0; k < 3; ++k) {
      if (!rn[k].empty()) ++nr;
      if (rn[k] == "PHOTON" || rn[k] == "CRPHOT" || rn[k] == "CRP") --nr;
    }
    for (int k = 0; k < 4; ++k) { if (!pn[k].empty()) ++np; if (pn[k] == "PHOTON") --np; }
    h.n_reac[i] = nr; h.n_prod[i] = np;
    // species are numbered by first appearance: reactants, then products
    for (int k = 0; k < nr; ++k) h.reac[3 * i + k] = species(rn[k]);
    for (int k = 0; k < np; ++k) h.prod[4 * i + k] = species(pn[k]);
  }
  const int N = h.N = (int)h.names.size();
  h.elements.assign(RACG_NELEM * N, 0); h.mass_num.assign(N, 0);
  h.vib_freq.assign(N, std::nan("")); h.Edesorb.assign(N, std::nan(""));
  for (int i = 0; i < N; ++i) {
    compose(h.names[i], &h.elements[RACG_NELEM * i]);
    double m = 0;
    for (int k = 0; k < RACG_NELEM; ++k) m += double(h.elements[RACG_NELEM * i + k]) * EMASS[k];
    h.mass_num[i] = m;
  }
  const double kB = 1.3806503e-16, mp = 1.67262158e-24, pi = 3.1415926535897932384626433;
  for (int i = 0; i < R; ++i)
    if (h.itype[i] == 62) {   // getVibFreq, src/chemistry.f90:1532-1539
      int r1 = h.reac[3 * i] - 1;
      double E = h.ABC[3 * i + 2];
      h.vib_freq[r1] = std::sqrt(2.0 * 1e15 * kB * E / (pi * pi) / (mp * h.mass_num[r1]));
      h.Edesorb[r1] = E;
    }
  // duplicate sets: earlier reactions with identical ctype, itype, reactants, products
  h.dupli_ptr.assign(R + 1, 0); h.dupli_list.clear();
  {
    std::unordered_map<std::string, std::vector<int>> seen;
    for (int i = 0; i < R; ++i) {
      std::string key((const char*)&h.reac[3 * i], 3 * sizeof(int));
      key.append((const char*)&h.prod[4 * i], 4 * sizeof(int));
      key.append((const char*)&h.itype[i], sizeof(int));
      key.push_back(h.ctype[2 * i]); key.push_back(h.ctype[2 * i + 1]);
      auto& v = seen[key];
      for (int j : v) h.dupli_list.push_back(j + 1);
      h.dupli_ptr[i + 1] = (int)h.dupli_list.size();
      v.push_back(i);
    }
  }
  return true;
}

}  // namespace

extern "C" {

typedef struct chem_host chem_host;

chem_host* chem_read_reactions(const char* filename_chemical_network, char* errbuf, int errlen) {
  ChemHost* h = new ChemHost();
  if (!read_network(*h, filename_chemical_network)) {
    if (errbuf && errlen > 0) { strncpy(errbuf, h->err.c_str(), errlen - 1); errbuf[errlen - 1] = 0; }
    delete h;
    return nullptr;
  }
  return (chem_host*)h;
}

void chem_host_free(chem_host* p) { delete (ChemHost*)p; }

void chem_host_sizes(const chem_host* p, int* R, int* N, int* ndupli) {
  const ChemHost* h = (const ChemHost*)p;
  *R = h->R; *N = h->N; *ndupli = (int)h->dupli_list.size();
}

// copies of chem_net%... / chem_species%... in the reference's layouts
void chem_host_tables(const chem_host* p, int* reac, int* prod, int* n_reac, int* n_prod, int* itype,
                      double* ABC, double* T_range, char* ctype, char* names, int* elements,
                      double* mass_num, double* vib_freq, double* Edesorb, int* dupli_ptr, int* dupli_list) {
  const ChemHost* h = (const ChemHost*)p;
  const int R = h->R, N = h->N;
  memcpy(reac, h->reac.data(), sizeof(int) * 3 * R); memcpy(prod, h->prod.data(), sizeof(int) * 4 * R);
  memcpy(n_reac, h->n_reac.data(), sizeof(int) * R); memcpy(n_prod, h->n_prod.data(), sizeof(int) * R);
  memcpy(itype, h->itype.data(), sizeof(int) * R);
  memcpy(ABC, h->ABC.data(), sizeof(double) * 3 * R); memcpy(T_range, h->T_range.data(), sizeof(double) * 2 * R);
  memcpy(ctype, h->ctype.data(), 2 * R);
  memset(names, ' ', (size_t)RACG_NAME_LEN * N);
  for (int i = 0; i < N; ++i) memcpy(names + (size_t)RACG_NAME_LEN * i, h->names[i].data(),
                                     std::min<size_t>(RACG_NAME_LEN, h->names[i].size()));
  memcpy(elements, h->elements.data(), sizeof(int) * RACG_NELEM * N);
  memcpy(mass_num, h->mass_num.data(), sizeof(double) * N);
  memcpy(vib_freq, h->vib_freq.data(), sizeof(double) * N);
  memcpy(Edesorb, h->Edesorb.data(), sizeof(double) * N);
  memcpy(dupli_ptr, h->dupli_ptr.data(), sizeof(int) * (R + 1));
  if (!h->dupli_list.empty()) memcpy(dupli_list, h->dupli_list.data(), sizeof(int) * h->dupli_list.size());
}

// y0(N): file values, neutralised with E-, renormalised to total H = 1
int chem_load_initial_abundances(const chem_host* p, const char* filename_initial_abundances, double* y0) {
  const ChemHost* h = (const ChemHost*)p;
  std::ifstream in(filename_initial_abundances);
  if (!in) return -1;
  std::unordered_map<std::string, int> id;
  for (int i = h->N - 1; i >= 0; --i) id[h->names[i]] = i;
  for (int i = 0; i < h->N; ++i) y0[i] = 0.0;
  std::string line;
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.size() < 64) line.append(64 - line.size(), ' ');
    auto it = id.find(strip(line.substr(0, RACG_NAME_LEN)));
    if (it != id.end()) y0[it->second] = field_real(line.substr(RACG_NAME_LEN, 16));
  }
  auto it = id.find("E-");
  if (it == id.end()) return -2;
  double q = 0.0, totH = 0.0;
  for (int i = 0; i < h->N; ++i) q += y0[i] * double(h->elements[RACG_NELEM * i]);
  y0[it->second] += q;
  if (y0[it->second] < 0.0) return -3;
  for (int i = 0; i < h->N; ++i) totH += double(h->elements[RACG_NELEM * i + 3]) * y0[i];
  for (int i = 0; i < h->N; ++i) y0[i] = y0[i] / totH;
  return 0;
}

}  // extern "C"
