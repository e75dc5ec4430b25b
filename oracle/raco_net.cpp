// raco_net.cpp -- ORACLE (test infrastructure only, see raco.h): network file ->
// tables, restating the reference's setup chain in src/chemistry.f90.
#include "raco_internal.hpp"
#include <fstream>
#include <cstdlib>
#include <algorithm>

namespace raco {

static thread_local std::string g_err;
void set_error(const std::string& s) { g_err = s; }
const std::string& get_error() { return g_err; }

static std::string rtrim(const std::string& s) {
  size_t e = s.size();
  while (e > 0 && s[e - 1] == ' ') --e;
  return s.substr(0, e);
}

// Fortran list-item read under edit descriptor Fw.0 from a fixed-width field:
// blanks are ignored (BLANK='NULL'), an all-blank field is 0, 'D'/'E' exponents
// and the letter-less form "1.5-3" are accepted, no decimal point => integer value
// (d = 0 so no implied scaling).  (format at src/chemistry.f90:1386-1387)
static double fortran_F(const std::string& field) {
  std::string s;
  for (char c : field) if (c != ' ') s.push_back(c);
  if (s.empty()) return 0.0;
  std::string m;  // mantissa / exponent split
  size_t i = 0;
  if (s[i] == '+' || s[i] == '-') m.push_back(s[i++]);
  while (i < s.size() && (isdigit((unsigned char)s[i]) || s[i] == '.')) m.push_back(s[i++]);
  int ex = 0;
  if (i < s.size()) {
    if (s[i] == 'D' || s[i] == 'd' || s[i] == 'E' || s[i] == 'e') ++i;
    ex = atoi(s.c_str() + i);
  }
  double v = m.empty() || m == "+" || m == "-" ? 0.0 : atof(m.c_str());
  if (ex != 0) v = atof((m + "e" + std::to_string(ex)).c_str());
  return v;
}

static int fortran_I(const std::string& field) {
  std::string s;
  for (char c : field) if (c != ' ') s.push_back(c);
  if (s.empty()) return 0;
  return atoi(s.c_str());
}

// getElements, src/chemistry.f90:1458-1529 (literal restatement; 1-based positions)
static const char* kElemNames[RACO_NELEM] = {"+-", "E", "Grain", "H", "D", "He", "C", "N", "O",
  "Si", "S", "Fe", "Na", "Mg", "Cl", "P", "F", "Ne", "Ar", "K"};
static const double kElemMass[RACO_NELEM] = {0.0, 5.45e-4, 0.0, 1.0, 2.0, 4.0, 12.0, 14.0, 16.0,
  28.0, 32.0, 56.0, 23.0, 24.0, 35.5, 31.0, 19.0, 20.18, 39.95, 39.1};

static void get_elements(const std::string& name_trim, int* arr) {
  for (int i = 0; i < RACO_NELEM; ++i) arr[i] = 0;
  // name padded to 12+ chars so (i+1) look-ahead reads a blank like the Fortran CHARACTER(12)
  std::string nm = name_trim;
  int lenName = (int)nm.size();
  nm.resize(34, ' ');
  int belongto[40] = {0};
  bool used[40] = {false};
  auto ch = [&](int p) { return nm[p - 1]; };  // 1-based
  for (int i = 1; i <= RACO_NELEM; ++i) {
    int lenEle = (int)strlen(kElemNames[i - 1]);
    for (int j = 1; j <= lenName - lenEle + 1; ++j) {
      if (nm.compare(j - 1, lenEle, kElemNames[i - 1]) == 0) {
        bool flagReplace = true;
        for (int k = j; k <= j + lenEle - 1; ++k) {
          if (used[k]) {
            if ((int)strlen(kElemNames[belongto[k] - 1]) >= lenEle) { flagReplace = false; break; }
            else arr[belongto[k] - 1] -= 1;
          }
        }
        if (flagReplace) {
          for (int k = j; k <= j + lenEle - 1; ++k) { belongto[k] = i; used[k] = true; }
          arr[i - 1] += 1;
        }
      }
    }
  }
  auto isdig = [](char c) { return c >= '0' && c <= '9'; };
  for (int i = 2; i <= lenName; ++i) {
    if (!used[i]) {
      for (int j = 1; j <= i - 1; ++j) {
        if (used[i - j]) { belongto[i] = belongto[i - j]; break; }
      }
      if (!isdig(ch(i - 1)) && isdig(ch(i))) {
        int ntmp;
        if (isdig(ch(i + 1))) ntmp = (ch(i) - '0') * 10 + (ch(i + 1) - '0');
        else ntmp = ch(i) - '0';
        if (ntmp == 0) continue;
        if (belongto[i] >= 1) arr[belongto[i] - 1] += ntmp - 1;
      } else if (ch(i) == '+') {
        arr[0] = 1;
      } else if (ch(i) == '-') {
        arr[0] = -1;
      }
    }
  }
}

// getVibFreq, src/chemistry.f90:1532-1539
static double get_vib_freq(double massnum, double Edesorb) {
  return std::sqrt(2.0 * const_SitesDensity_CGS * phy_kBoltzmann_CGS * Edesorb / (phy_Pi * phy_Pi) /
                   (phy_mProton_CGS * massnum));
}

static bool load(Net& n, const char* file) {
  std::ifstream in(file);
  if (!in) { set_error(std::string("cannot open ") + file); return false; }
  // chem_read_reactions, src/chemistry.f90:1427-1454: keep rows whose first
  // character is neither '!' nor blank.
  std::vector<std::string> rows;
  std::string line;
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty()) continue;
    if (line[0] == '!' || line[0] == ' ') continue;
    line.resize(150, ' ');  // const_len_reactionfile_row
    rows.push_back(line);
  }
  n.R = (int)rows.size();
  const int R = n.R;
  n.reac_names.resize(R); n.prod_names.resize(R);
  n.reac.assign(3 * R, 0); n.prod.assign(4 * R, 0);
  n.n_reac.assign(R, 0); n.n_prod.assign(R, 0); n.itype.assign(R, 0);
  n.ABC.assign(3 * R, 0.0); n.T_range.assign(2 * R, 0.0);
  n.ctype.resize(R);
  // chem_load_reactions, src/chemistry.f90:1364-1424
  // FMT (7(A12), 3F9.0, 2F6.0, I3, X, A1, X, A2)
  for (int i = 0; i < R; ++i) {
    const std::string& s = rows[i];
    for (int k = 0; k < 3; ++k) n.reac_names[i][k] = s.substr(12 * k, 12);
    for (int k = 0; k < 4; ++k) n.prod_names[i][k] = s.substr(36 + 12 * k, 12);
    for (int k = 0; k < 3; ++k) n.ABC[3 * i + k] = fortran_F(s.substr(84 + 9 * k, 9));
    for (int k = 0; k < 2; ++k) n.T_range[2 * i + k] = fortran_F(s.substr(111 + 6 * k, 6));
    n.itype[i] = fortran_I(s.substr(123, 3));
    n.ctype[i] = s.substr(129, 2);
    for (int j = 0; j < 3; ++j) {
      std::string t = rtrim(n.reac_names[i][j]);
      bool nonblank = n.reac_names[i][j].find_first_not_of(' ') != std::string::npos;
      if (nonblank) n.n_reac[i] += 1;
      if (t == "PHOTON") n.n_reac[i] -= 1;
      if (t == "CRPHOT") n.n_reac[i] -= 1;
      if (t == "CRP") n.n_reac[i] -= 1;
    }
    for (int j = 0; j < 4; ++j) {
      std::string t = rtrim(n.prod_names[i][j]);
      bool nonblank = n.prod_names[i][j].find_first_not_of(' ') != std::string::npos;
      if (nonblank) n.n_prod[i] += 1;
      if (t == "PHOTON") n.n_prod[i] -= 1;
    }
  }
  // chem_parse_reactions, src/chemistry.f90:1221-1360: species numbered by first
  // appearance, reactants then products, reaction by reaction (linear search).
  std::vector<std::string>& names = n.names;
  names.clear();
  names.push_back(rtrim(n.reac_names[0][0]));
  auto lookup = [&](const std::string& nm) -> int {
    for (size_t j = 0; j < names.size(); ++j) if (names[j] == nm) return (int)j + 1;
    names.push_back(nm);
    return (int)names.size();
  };
  for (int i = 0; i < R; ++i) {
    for (int k = 0; k < n.n_reac[i]; ++k) n.reac[3 * i + k] = lookup(rtrim(n.reac_names[i][k]));
    for (int k = 0; k < n.n_prod[i]; ++k) n.prod[4 * i + k] = lookup(rtrim(n.prod_names[i][k]));
  }
  n.N = (int)names.size();
  n.NEQ = n.N + 1;
  const int N = n.N;
  n.elements.assign(RACO_NELEM * N, 0);
  n.mass_num.assign(N, 0.0);
  const double qnan = std::nan("");
  n.vib_freq.assign(N, qnan); n.Edesorb.assign(N, qnan);
  n.counterpart.assign(N, -1);
  for (int i = 0; i < N; ++i) {
    get_elements(names[i], &n.elements[RACO_NELEM * i]);
    double m = 0.0;
    for (int e = 0; e < RACO_NELEM; ++e) m += double(n.elements[RACO_NELEM * i + e]) * kElemMass[e];
    n.mass_num[i] = m;
  }
  for (int i = 0; i < R; ++i) {
    if (n.itype[i] == 62) {  // src/chemistry.f90:1322-1331
      int r1 = n.reac[3 * i], p1 = n.prod[4 * i];
      n.vib_freq[r1 - 1] = get_vib_freq(n.mass_num[r1 - 1], n.ABC[3 * i + 2]);
      n.Edesorb[r1 - 1] = n.ABC[3 * i + 2];
      n.counterpart[p1 - 1] = r1;
      n.counterpart[r1 - 1] = p1;
    }
  }
  n.grain_idx.clear();
  for (int i = 0; i < N; ++i) if (names[i][0] == 'g') n.grain_idx.push_back(i + 1);
  n.nGrain = (int)n.grain_idx.size();
  // chem_get_dupli_reactions, src/chemistry.f90:1188-1217
  n.dupli.assign(R, {});
  for (int i = 0; i < R; ++i) {
    for (int j = 0; j < i; ++j) {
      if (n.itype[i] != n.itype[j] || n.ctype[i] != n.ctype[j]) continue;
      bool same = true;
      for (int k = 0; k < 3 && same; ++k) same = n.reac[3 * i + k] == n.reac[3 * j + k];
      for (int k = 0; k < 4 && same; ++k) same = n.prod[4 * i + k] == n.prod[4 * j + k];
      if (same) n.dupli[i].push_back(j + 1);
    }
  }
  // chem_get_idx_for_special_species, src/chemistry.f90:1089-1185
  static const struct { const char* nm; int pos; } sp[] = {
    {"H2", S_H2}, {"H", S_HI}, {"E-", S_E}, {"C", S_CI}, {"C+", S_CII}, {"O", S_OI}, {"O2", S_O2},
    {"CO", S_CO}, {"H2O", S_H2O}, {"OH", S_OH}, {"H+", S_Hplus}, {"He+", S_Heplus}, {"gH", S_gH},
    {"gH2", S_gH2}, {"Grain0", S_Grain0}, {"Grain-", S_GrainM}, {"Grain+", S_GrainP},
    {"gH2O", S_gH2O}, {"gCO", S_gCO}, {"gCO2", S_gCO2}, {"gN2", S_gN2}, {"N+", S_NII},
    {"Si+", S_SiII}, {"Fe+", S_FeII}, {"N", S_NI}};
  for (int k = 0; k < 32; ++k) n.special[k] = 0;
  for (int i = 0; i < N; ++i)
    for (auto& e : sp) if (names[i] == e.nm) n.special[e.pos] = i + 1;
  // per-reaction string predicates pre-resolved
  n.first_is_H2.assign(R, 0); n.first_is_gH.assign(R, 0); n.fss_kind.assign(R, 0);
  for (int i = 0; i < R; ++i) {
    std::string r1 = rtrim(n.reac_names[i][0]);
    n.first_is_H2[i] = (r1 == "H2");
    n.first_is_gH[i] = (r1 == "gH");
    // f_selfshielding_toISM/toStar, src/chemistry.f90:1007-1063: keyed on the NAME of
    // species reac(1,i) and only for ctype PH / LA
    int kind = 0;
    if ((n.ctype[i] == "PH" || n.ctype[i] == "LA") && n.reac[3 * i] > 0) {
      const std::string& sn = names[n.reac[3 * i] - 1];
      if (sn == "H2") kind = 1; else if (sn == "CO") kind = 2;
      else if (sn == "H2O") kind = 3; else if (sn == "OH") kind = 4;
    }
    n.fss_kind[i] = (char)kind;
  }
  // chem_make_sparse_structure, src/chemistry.f90:1858-1885
  const int NEQ = n.NEQ;
  std::vector<char> mask((size_t)NEQ * NEQ, 0);  // mask[row + col*NEQ]
  for (int i = 0; i < R; ++i) {
    for (int j = 0; j < n.n_reac[i]; ++j) {
      int col = n.reac[3 * i + j] - 1;
      for (int k = 0; k < n.n_reac[i]; ++k) mask[(size_t)(n.reac[3 * i + k] - 1) + (size_t)col * NEQ] = 1;
      for (int k = 0; k < n.n_prod[i]; ++k) mask[(size_t)(n.prod[4 * i + k] - 1) + (size_t)col * NEQ] = 1;
    }
  }
  for (int i = 0; i < NEQ; ++i) mask[(size_t)i + (size_t)(NEQ - 1) * NEQ] = 1;
  for (int i = 0; i < 10; ++i)
    if (n.special[i] > 0) mask[(size_t)(NEQ - 1) + (size_t)(n.special[i] - 1) * NEQ] = 1;
  // chem_prepare_solver_storage, src/chemistry.f90:1962-1971 (IA/JA 1-based, column-major)
  n.ia.assign(NEQ + 1, 0); n.ja.clear();
  n.slot_of.assign((size_t)NEQ * NEQ, -1);
  n.ia[0] = 1;
  int k = 1;
  for (int c = 0; c < NEQ; ++c) {
    for (int r = 0; r < NEQ; ++r) {
      if (mask[(size_t)r + (size_t)c * NEQ]) {
        n.slot_of[(size_t)r + (size_t)c * NEQ] = k - 1;
        n.ja.push_back(r + 1);
        ++k;
      }
    }
    n.ia[c + 1] = k;
  }
  n.NNZ = (int)n.ja.size();
  return true;
}

}  // namespace raco

using namespace raco;

extern "C" {

const char* raco_last_error(void) { return raco::get_error().c_str(); }

raco_net* raco_net_load(const char* file) {
  raco_net* h = new raco_net();
  if (!load(h->net, file)) { delete h; return nullptr; }
  return h;
}
void raco_net_free(raco_net* h) { delete h; }

void raco_net_sizes(const raco_net* h, int* s) {
  const Net& n = h->net;
  int nd = 0;
  for (auto& d : n.dupli) nd += (int)d.size();
  s[0] = n.R; s[1] = n.N; s[2] = n.NEQ; s[3] = n.NNZ; s[4] = n.nGrain; s[5] = nd;
  SparseLU lu;
  lu.analyse(n.NEQ, n.ia, n.ja);
  s[6] = lu.nnz_a; s[7] = lu.nzl + lu.nzu + n.NEQ;
}

void raco_net_tables(const raco_net* h, int* reac, int* prod, int* n_reac, int* n_prod, int* itype,
                     double* ABC, double* T_range, char* ctype) {
  const Net& n = h->net;
  if (reac) memcpy(reac, n.reac.data(), sizeof(int) * 3 * n.R);
  if (prod) memcpy(prod, n.prod.data(), sizeof(int) * 4 * n.R);
  if (n_reac) memcpy(n_reac, n.n_reac.data(), sizeof(int) * n.R);
  if (n_prod) memcpy(n_prod, n.n_prod.data(), sizeof(int) * n.R);
  if (itype) memcpy(itype, n.itype.data(), sizeof(int) * n.R);
  if (ABC) memcpy(ABC, n.ABC.data(), sizeof(double) * 3 * n.R);
  if (T_range) memcpy(T_range, n.T_range.data(), sizeof(double) * 2 * n.R);
  if (ctype) for (int i = 0; i < n.R; ++i) { ctype[2 * i] = n.ctype[i][0]; ctype[2 * i + 1] = n.ctype[i][1]; }
}

void raco_net_species(const raco_net* h, char* names, int* elements, double* mass_num,
                      double* vib_freq, double* Edesorb, int* idx_counterpart) {
  const Net& n = h->net;
  if (names) {
    memset(names, ' ', (size_t)RACO_NAME_LEN * n.N);
    for (int i = 0; i < n.N; ++i) memcpy(names + (size_t)RACO_NAME_LEN * i, n.names[i].data(),
                                         std::min<size_t>(RACO_NAME_LEN, n.names[i].size()));
  }
  if (elements) memcpy(elements, n.elements.data(), sizeof(int) * RACO_NELEM * n.N);
  if (mass_num) memcpy(mass_num, n.mass_num.data(), sizeof(double) * n.N);
  if (vib_freq) memcpy(vib_freq, n.vib_freq.data(), sizeof(double) * n.N);
  if (Edesorb) memcpy(Edesorb, n.Edesorb.data(), sizeof(double) * n.N);
  if (idx_counterpart) memcpy(idx_counterpart, n.counterpart.data(), sizeof(int) * n.N);
}

void raco_net_dupli(const raco_net* h, int* ptr, int* list) {
  const Net& n = h->net;
  int k = 0;
  for (int i = 0; i < n.R; ++i) {
    ptr[i] = k;
    for (int j : n.dupli[i]) list[k++] = j;
  }
  ptr[n.R] = k;
}

void raco_net_special(const raco_net* h, int* special) { memcpy(special, h->net.special, sizeof(int) * 32); }

void raco_net_grain_species(const raco_net* h, int* idx) {
  memcpy(idx, h->net.grain_idx.data(), sizeof(int) * h->net.nGrain);
}

void raco_net_pattern(const raco_net* h, int* ia, int* ja) {
  const Net& n = h->net;
  memcpy(ia, n.ia.data(), sizeof(int) * (n.NEQ + 1));
  memcpy(ja, n.ja.data(), sizeof(int) * n.NNZ);
}

// chem_load_initial_abundances, src/chemistry.f90:1978-2024
int raco_load_initial_abundances(const raco_net* h, const char* file, double* y0) {
  const Net& n = h->net;
  std::ifstream in(file);
  if (!in) { set_error(std::string("cannot open ") + file); return -1; }
  for (int i = 0; i < n.N; ++i) y0[i] = 0.0;
  std::string line;
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    line.resize(64, ' ');  // const_len_init_abun_file_row
    std::string nm = raco::rtrim(line.substr(0, RACO_NAME_LEN));
    for (int i = 0; i < n.N; ++i) {
      if (nm == n.names[i]) { y0[i] = raco::fortran_F(line.substr(RACO_NAME_LEN, 16)); break; }
    }
  }
  int iE = n.special[S_E];
  if (iE <= 0) { set_error("no E- species"); return -2; }
  double q = 0.0;
  for (int i = 0; i < n.N; ++i) q += y0[i] * double(n.elements[RACO_NELEM * i + 0]);
  y0[iE - 1] += q;
  if (y0[iE - 1] < 0.0) { set_error("Cannot neutralize the initial condition!"); return -3; }
  double totH = 0.0;
  for (int i = 0; i < n.N; ++i) totH += double(n.elements[RACO_NELEM * i + 3]) * y0[i];
  for (int i = 0; i < n.N; ++i) y0[i] = y0[i] / totH;
  return 0;
}

}  // extern "C"
