// racg_host.cpp -- host-side network setup of libracg (see racg_host.hpp).
#include "racg_host.hpp"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <numeric>

namespace racg {

static std::string trim_name(const char* p, int len) {
  int e = len;
  while (e > 0 && (p[e - 1] == ' ' || p[e - 1] == '\0')) --e;
  return std::string(p, p + e);
}

// Cut rows of (idx, coef) pairs into sub-rows of <= SEG entries, sort sub-rows by
// length (longest first) and pack 32 of them per block, transposed.
static void build_gather(Gather& g, const std::vector<std::vector<std::pair<int, int>>>& rows,
                         int SEG) {
  struct Sub { int row, first, len, target; };
  std::vector<Sub> subs;
  g.nrows = (int)rows.size();
  g.npartial = 0;
  g.comb_row.clear(); g.comb_ptr.assign(1, 0);
  g.nent_real = 0;
  for (int r = 0; r < (int)rows.size(); ++r) {
    int len = (int)rows[r].size();
    g.nent_real += len;
    if (len == 0) continue;
    int nsub = (len + SEG - 1) / SEG;
    if (nsub == 1) subs.push_back({r, 0, len, r});
    else {
      g.comb_row.push_back(r);
      for (int s = 0; s < nsub; ++s) {
        int f = s * SEG;
        subs.push_back({r, f, std::min(SEG, len - f), -2 - g.npartial});
        ++g.npartial;
      }
      g.comb_ptr.push_back(g.npartial);
    }
  }
  g.ncombine = (int)g.comb_row.size();
  std::stable_sort(subs.begin(), subs.end(), [](const Sub& a, const Sub& b) { return a.len > b.len; });
  g.nblk = ((int)subs.size() + 31) / 32;
  g.blk_off.assign(g.nblk + 1, 0); g.blk_width.assign(g.nblk, 0);
  g.sub_target.assign((size_t)g.nblk * 32, -1);
  g.ent.clear();
  for (int b = 0; b < g.nblk; ++b) {
    int w = 0;
    for (int l = 0; l < 32; ++l) {
      size_t s = (size_t)b * 32 + l;
      if (s < subs.size()) { w = std::max(w, subs[s].len); g.sub_target[s] = subs[s].target; }
    }
    g.blk_width[b] = w;
    g.blk_off[b] = (int)g.ent.size();
    g.ent.resize(g.ent.size() + (size_t)w * 32, (uint32_t)(0u | (4u << 24)));  // idx 0, coef 0
    for (int l = 0; l < 32; ++l) {
      size_t s = (size_t)b * 32 + l;
      if (s >= subs.size()) continue;
      const Sub& sb = subs[s];
      for (int j = 0; j < sb.len; ++j) {
        auto& e = rows[sb.row][sb.first + j];
        g.ent[(size_t)g.blk_off[b] + (size_t)j * 32 + l] = (uint32_t)e.first | ((uint32_t)(e.second + 4) << 24);
      }
    }
  }
  g.blk_off[g.nblk] = (int)g.ent.size();
}

bool build_host_net(HostNet& hn, int R, int N, const int* reac, const int* prod, const int* n_reac,
                    const int* n_prod, const int* itype, const double* ABC, const double* T_range,
                    const char* ctype, const char* names, const int* elements,
                    const double* mass_num, const double* vib_freq, const double* Edesorb,
                    const int* dupli_ptr, const int* dupli_list, const racg_cfg* cfg) {
  if (R <= 0 || N <= 0 || N > 1000) { hn.error = "bad R/N (N must be <= 1000)"; return false; }
  hn.R = R; hn.N = N; hn.NEQ = N + 1; hn.n = N;
  hn.cfg = *cfg;
  if (cfg->H2_form_use_moeq || cfg->update_gH_params_realtime || cfg->evol_dust_size) {
    hn.error = "H2_form_use_moeq / update_gH_params_realtime / evol_dust_size = .true. are not supported";
    return false;
  }
  hn.reac.assign(reac, reac + 3 * R); hn.prod.assign(prod, prod + 4 * R);
  hn.n_reac.assign(n_reac, n_reac + R); hn.n_prod.assign(n_prod, n_prod + R);
  hn.itype.assign(itype, itype + R);
  hn.ABC.assign(ABC, ABC + 3 * R); hn.T_range.assign(T_range, T_range + 2 * R);
  hn.elements.assign(elements, elements + RACG_NELEM * N);
  hn.mass_num.assign(mass_num, mass_num + N); hn.vib_freq.assign(vib_freq, vib_freq + N);
  hn.Edesorb.assign(Edesorb, Edesorb + N);
  hn.dupli_ptr.assign(dupli_ptr, dupli_ptr + R + 1);
  hn.dupli_list.assign(dupli_list, dupli_list + dupli_ptr[R]);
  hn.names.resize(N); hn.ctype.resize(R);
  for (int i = 0; i < N; ++i) hn.names[i] = trim_name(names + (size_t)RACG_NAME_LEN * i, RACG_NAME_LEN);
  for (int i = 0; i < R; ++i) hn.ctype[i] = std::string(ctype + 2 * i, ctype + 2 * i + 2);
  for (int i = 0; i < R; ++i) {
    if (hn.n_reac[i] < 1 || hn.n_reac[i] > 2 || hn.n_prod[i] < 0 || hn.n_prod[i] > 4) {
      hn.error = "reaction " + std::to_string(i + 1) + ": n_reac must be 1..2 and n_prod 0..4";
      return false;
    }
    for (int k = 0; k < hn.n_reac[i]; ++k)
      if (hn.reac[3 * i + k] < 1 || hn.reac[3 * i + k] > N) { hn.error = "reactant id out of range"; return false; }
    for (int k = 0; k < hn.n_prod[i]; ++k)
      if (hn.prod[4 * i + k] < 1 || hn.prod[4 * i + k] > N) { hn.error = "product id out of range"; return false; }
  }
  // special species by name (chem_get_idx_for_special_species, src/chemistry.f90:1089-1185)
  std::map<std::string, int> byname;
  for (int i = N - 1; i >= 0; --i) byname[hn.names[i]] = i;
  auto find = [&](const char* s) { auto it = byname.find(s); return it == byname.end() ? -1 : it->second; };
  hn.iH2 = find("H2"); hn.iH = find("H"); hn.iE = find("E-"); hn.igH = find("gH"); hn.igH2 = find("gH2");
  hn.igH2O = find("gH2O"); hn.iGrain0 = find("Grain0"); hn.iGrainM = find("Grain-"); hn.iGrainP = find("Grain+");
  hn.hc_idx.clear();
  for (const char* s : {"H2", "H", "E-", "C", "C+", "O", "O2", "CO", "H2O", "OH"}) hn.hc_idx.push_back(find(s));
  hn.grain_idx.clear();
  for (int i = 0; i < N; ++i) if (!hn.names[i].empty() && hn.names[i][0] == 'g') hn.grain_idx.push_back(i);

  // ---- Jacobian pattern, bit-exact with the reference (src/chemistry.f90:1866-1884, 1962-1971)
  const int NEQ = hn.NEQ;
  {
    std::vector<uint8_t> mask((size_t)NEQ * NEQ, 0);  // [col*NEQ + row]
    for (int i = 0; i < R; ++i)
      for (int j = 0; j < hn.n_reac[i]; ++j) {
        size_t c = (size_t)(hn.reac[3 * i + j] - 1) * NEQ;
        for (int k = 0; k < hn.n_reac[i]; ++k) mask[c + hn.reac[3 * i + k] - 1] = 1;
        for (int k = 0; k < hn.n_prod[i]; ++k) mask[c + hn.prod[4 * i + k] - 1] = 1;
      }
    for (int i = 0; i < NEQ; ++i) mask[(size_t)(NEQ - 1) * NEQ + i] = 1;
    for (int s : hn.hc_idx) if (s >= 0) mask[(size_t)s * NEQ + (NEQ - 1)] = 1;
    hn.ia.assign(NEQ + 1, 0); hn.ja.clear();
    hn.ia[0] = 1;
    int ndiag_missing = 0;
    for (int c = 0; c < NEQ; ++c) {
      for (int r = 0; r < NEQ; ++r) if (mask[(size_t)c * NEQ + r]) hn.ja.push_back(r + 1);
      hn.ia[c + 1] = (int)hn.ja.size() + 1;
      if (!mask[(size_t)c * NEQ + c]) ++ndiag_missing;
    }
    hn.NNZ = (int)hn.ja.size();
    hn.NNZ_diag = hn.NNZ + ndiag_missing;
  }

  // ---- rate tables (branches of chem_cal_rates, src/chemistry.f90:680-934)
  hn.rcode.assign(R, 0);
  hn.rA.assign(R, 0); hn.rB.assign(R, 0); hn.rC.assign(R, 0); hn.rTlo.assign(R, 0); hn.rThi.assign(R, 0);
  hn.rX.assign((size_t)6 * R, 0.0);
  for (int i = 0; i < R; ++i) {
    hn.rA[i] = ABC[3 * i]; hn.rB[i] = ABC[3 * i + 1]; hn.rC[i] = ABC[3 * i + 2];
    hn.rTlo[i] = T_range[2 * i]; hn.rThi[i] = T_range[2 * i + 1];
    const int r1 = hn.reac[3 * i] - 1, r2 = hn.n_reac[i] >= 2 ? hn.reac[3 * i + 1] - 1 : -1;
    int cls = RC_ZERO, fss = 0, sigcheck = 0;
    if ((hn.ctype[i] == "PH" || hn.ctype[i] == "LA")) {  // src/chemistry.f90:1007-1063
      const std::string& s = hn.names[r1];
      fss = s == "H2" ? 1 : s == "CO" ? 2 : s == "H2O" ? 3 : s == "OH" ? 4 : 0;
    }
    double* X = &hn.rX[(size_t)6 * i];
    switch (hn.itype[i]) {
      case 5: cls = RC_ARRH; break;
      case 6: cls = RC_ARRH_STRICT; break;
      case 1: cls = RC_CR; break;
      case 2: case 20: cls = RC_CRPHOT; break;
      case 3: cls = (hn.names[r1] == "H2") ? RC_PHOTO_H2 : RC_PHOTO; break;
      case 13: cls = RC_LYA; break;
      case 21: {
        if (r2 < 0) { hn.error = "type 21 reaction needs two reactants"; return false; }
        int id3;
        if (elements[RACG_NELEM * r1 + 2] == 0) id3 = r1;
        else if (elements[RACG_NELEM * r2 + 2] == 0) id3 = r2;
        else { hn.error = "Species name problem with type 21."; return false; }
        int q = elements[RACG_NELEM * r1] * elements[RACG_NELEM * r2];
        if (q == -1) cls = RC_GRAIN_NP; else if (q == 0) cls = RC_GRAIN_N0;
        else { hn.error = "Charge problem with type 21."; return false; }
        X[0] = mass_num[id3];
        break;
      }
      case 0: cls = RC_H2FORM0; X[0] = mass_num[r1]; break;
      case 61: cls = RC_ADSORB; X[0] = mass_num[r1]; break;
      case 62: cls = RC_DESORB; X[0] = vib_freq[r1]; break;
      case 63: cls = RC_SURF_AA; X[0] = vib_freq[r1]; X[1] = mass_num[r1]; X[2] = Edesorb[r1];
        sigcheck = (hn.names[r1] == "gH"); break;
      case 64:
        if (r2 < 0) { hn.error = "type 64 reaction needs two reactants"; return false; }
        cls = RC_SURF_AB; X[0] = vib_freq[r1]; X[1] = mass_num[r1]; X[2] = Edesorb[r1];
        X[3] = vib_freq[r2]; X[4] = mass_num[r2]; X[5] = Edesorb[r2]; break;
      case 75: cls = RC_PHOTODES; break;
      default: cls = RC_ZERO;
    }
    int two_body_gas = (hn.n_reac[i] == 2 && hn.itype[i] < 60) ? 1 : 0;
    hn.rcode[i] = cls | (fss << 8) | (two_body_gas << 12) | (sigcheck << 13);
  }

  // ---- flux words + net stoichiometric coefficients (src/disk.f90:4583-4650)
  hn.fw.assign(R, 0); hn.sat_c.clear();
  std::vector<std::vector<std::pair<int, int>>> sp_rows(N);   // species -> (reaction, coef)
  hn.rx_species.assign((size_t)6 * R, -1);
  for (int i = 0; i < R; ++i) {
    int kind;
    switch (hn.itype[i]) {
      case 5: case 6: case 21: case 64: kind = FK_TWO; break;
      case 1: case 2: case 3: case 13: case 61: case 20: case 0: kind = FK_ONE; break;
      case 62: case 75: kind = FK_SAT; break;
      case 63: kind = FK_TWO; break;   // k*y1*y1, sign flip when y1<0 == two-body with r2 = r1
      default: kind = FK_SKIP;
    }
    int r1 = hn.reac[3 * i] - 1;
    int r2 = (hn.n_reac[i] >= 2) ? hn.reac[3 * i + 1] - 1 : -1;
    if (hn.itype[i] == 63) r2 = r1;
    if (kind == FK_TWO && r2 < 0) { hn.error = "two-body reaction type with one reactant"; return false; }
    uint32_t sat = 0;
    if (kind == FK_SAT) {
      sat = (uint32_t)hn.sat_c.size();
      hn.sat_c.push_back(hn.itype[i] == 75 ? hn.rC[i] : 1.0);
      if (sat > 1023) { hn.error = "too many saturating reactions"; return false; }
    }
    hn.fw[i] = (uint32_t)r1 | ((uint32_t)(kind == FK_TWO ? r2 : 1023) << 10) | ((uint32_t)kind << 20) | (sat << 22);
    hn.rx_species[(size_t)6 * i + 0] = r1;
    hn.rx_species[(size_t)6 * i + 1] = (kind == FK_TWO) ? r2 : -1;
    for (int k = 0; k < hn.n_prod[i]; ++k) hn.rx_species[(size_t)6 * i + 2 + k] = hn.prod[4 * i + k] - 1;
    if (kind == FK_SKIP) continue;
    std::map<int, int> net;
    for (int k = 0; k < hn.n_reac[i]; ++k) net[hn.reac[3 * i + k] - 1] -= 1;
    for (int k = 0; k < hn.n_prod[i]; ++k) net[hn.prod[4 * i + k] - 1] += 1;
    for (auto& kv : net) {
      if (kv.second == 0) continue;
      if (kv.second < -4 || kv.second > 3) { hn.error = "stoichiometric coefficient out of range"; return false; }
      sp_rows[kv.first].push_back({i, kv.second});
    }
  }
  hn.nsat = (int)hn.sat_c.size();
  build_gather(hn.rhs, sp_rows, 32);

  // ---- species-block pattern and fill-reducing ordering (minimum degree on M + M^T)
  const int n = N;
  const int W = (n + 63) / 64;
  std::vector<uint64_t> A((size_t)n * W, 0);   // A[row][col] bits, unsymmetric pattern incl. diagonal
  auto setb = [&](std::vector<uint64_t>& M, int r, int c) { M[(size_t)r * W + (c >> 6)] |= (1ull << (c & 63)); };
  auto getb = [&](const std::vector<uint64_t>& M, int r, int c) { return (M[(size_t)r * W + (c >> 6)] >> (c & 63)) & 1ull; };
  for (int c = 0; c < n; ++c)
    for (int k = hn.ia[c] - 1; k < hn.ia[c + 1] - 1; ++k) { int r = hn.ja[k] - 1; if (r < n) setb(A, r, c); }
  for (int i = 0; i < n; ++i) setb(A, i, i);
  std::vector<uint64_t> G((size_t)n * W, 0);   // symmetric elimination graph (no self loops)
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) if (r != c && (getb(A, r, c) || getb(A, c, r))) { setb(G, r, c); setb(G, c, r); }
  hn.perm.assign(n, 0); hn.iperm.assign(n, 0);
  {
    std::vector<char> done(n, 0);
    std::vector<int> deg(n);
    auto degree = [&](int v) { int d = 0; for (int w = 0; w < W; ++w) d += __builtin_popcountll(G[(size_t)v * W + w]); return d; };
    for (int v = 0; v < n; ++v) deg[v] = degree(v);
    int clique_start = -1;
    for (int step = 0; step < n; ++step) {
      int best = -1;
      for (int v = 0; v < n; ++v) if (!done[v] && (best < 0 || deg[v] < deg[best])) best = v;
      if (clique_start < 0 && deg[best] >= n - step - 1) clique_start = step;
      done[best] = 1; hn.perm[step] = best; hn.iperm[best] = step;
      // neighbours of best become a clique
      std::vector<uint64_t> nb(G.begin() + (size_t)best * W, G.begin() + (size_t)best * W + W);
      for (int a = 0; a < n; ++a) {
        if (!((nb[a >> 6] >> (a & 63)) & 1ull)) continue;
        uint64_t* ga = &G[(size_t)a * W];
        for (int w = 0; w < W; ++w) ga[w] |= nb[w];
        ga[a >> 6] &= ~(1ull << (a & 63));
        ga[best >> 6] &= ~(1ull << (best & 63));
        deg[a] = degree(a);
      }
      for (int w = 0; w < W; ++w) G[(size_t)best * W + w] = 0;
    }
    if (clique_start < 0) clique_start = n - 1;
    int nt = n - clique_start;
    nt = ((nt + 15) / 16) * 16;
    nt = std::max(16, std::min(nt, n));
    // the Schur complement (nt*nt) plus one dense work row per warp (8 warps) must fit
    // the integrator's shared memory next to its 12 length-n vectors (racg_integrate.cu)
    {
      const double budget = 227.0 * 1024 - 8.0 * (12.0 * n + 64) - 8.0 * 8 * n - 2048;
      int cap = (int)std::floor(std::sqrt(std::max(budget, 0.0) / 8.0));
      cap = (cap / 16) * 16;
      if (cap < 16) { hn.error = "network too large for the integrator's shared-memory layout"; return false; }
      nt = std::min(nt, cap);
    }
    hn.nt = nt; hn.nh = n - nt;
  }
  // ---- symbolic LU on the permuted pattern (row merge; no pivoting)
  std::vector<uint64_t> F((size_t)n * W, 0);
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) if (getb(A, hn.perm[r], hn.perm[c])) setb(F, r, c);
  for (int i = 0; i < n; ++i) {
    uint64_t* fi = &F[(size_t)i * W];
    for (int k = 0; k < i; ++k) {
      if (!((fi[k >> 6] >> (k & 63)) & 1ull)) continue;
      const uint64_t* fk = &F[(size_t)k * W];
      // merge the U part of row k (cols > k)
      for (int w = (k >> 6); w < W; ++w) {
        uint64_t m = fk[w];
        if (w == (k >> 6)) m &= ~((2ull << (k & 63)) - 1ull);
        fi[w] |= m;
      }
    }
  }
  const int nh = hn.nh, nt = hn.nt;
  hn.nnz_lu = 0;
  for (int i = 0; i < n; ++i) for (int w = 0; w < W; ++w) hn.nnz_lu += __builtin_popcountll(F[(size_t)i * W + w]);
  hn.row_ptr.assign(n + 1, 0); hn.row_nl.assign(n, 0); hn.col.clear();
  for (int i = 0; i < n; ++i) {
    int nl = 0;
    int cmax = (i < nh) ? n : nh;   // tail rows keep only the L_C part (cols < nh) in CSR
    for (int c = 0; c < cmax; ++c) {
      if (!getb(F, i, c) && c != i) continue;
      hn.col.push_back((uint16_t)c);
      if (c < i) ++nl;
    }
    hn.row_nl[i] = nl;
    hn.row_ptr[i + 1] = (int)hn.col.size();
  }
  hn.nslots = (int)hn.col.size();
  // levels
  auto make_levels = [&](const std::vector<int>& lev, int count, std::vector<int>& ptr, std::vector<int>& rows) {
    int nlev = 0;
    for (int i = 0; i < count; ++i) nlev = std::max(nlev, lev[i] + 1);
    ptr.assign(nlev + 1, 0);
    for (int i = 0; i < count; ++i) ptr[lev[i] + 1]++;
    for (int l = 0; l < nlev; ++l) ptr[l + 1] += ptr[l];
    rows.assign(count, 0);
    std::vector<int> pos(ptr.begin(), ptr.end() - 1);
    for (int i = 0; i < count; ++i) rows[pos[lev[i]]++] = i;
  };
  {
    std::vector<int> lev(nh, 0);
    for (int i = 0; i < nh; ++i) {
      int l = 0;
      for (int q = hn.row_ptr[i]; q < hn.row_ptr[i] + hn.row_nl[i]; ++q) l = std::max(l, lev[hn.col[q]] + 1);
      lev[i] = l;
    }
    make_levels(lev, nh, hn.flev_ptr, hn.flev_rows);
    hn.sl_ptr = hn.flev_ptr; hn.sl_rows = hn.flev_rows;
    std::vector<int> levu(nh, 0);
    for (int i = nh - 1; i >= 0; --i) {
      int l = 0;
      for (int q = hn.row_ptr[i] + hn.row_nl[i] + 1; q < hn.row_ptr[i + 1]; ++q) {
        int c = hn.col[q];
        if (c < nh) l = std::max(l, levu[c] + 1);
      }
      levu[i] = l;
    }
    make_levels(levu, nh, hn.su_ptr, hn.su_rows);
  }

  // ---- Jacobian gather: J(i,j) = sum_r coef_i(r) * dflux[r][which(j)]  (src/disk.f90:4765-4875)
  // storage index of (pi,pj) in permuted space
  auto store_index = [&](int pi, int pj) -> int {
    if (pi >= nh && pj >= nh) return hn.nslots + (pi - nh) * nt + (pj - nh);
    int lo = hn.row_ptr[pi], hi = hn.row_ptr[pi + 1];
    auto it = std::lower_bound(hn.col.begin() + lo, hn.col.begin() + hi, (uint16_t)pj);
    if (it == hn.col.begin() + hi || *it != pj) return -1;
    return (int)(it - hn.col.begin());
  };
  {
    std::map<int, std::vector<std::pair<int, int>>> tgt;   // store idx -> (r*2+which, coef)
    for (int i = 0; i < R; ++i) {
      uint32_t w = hn.fw[i];
      int kind = (w >> 20) & 3;
      if (kind == FK_SKIP) continue;
      int r1 = w & 1023, r2 = (w >> 10) & 1023;
      int ncol = (kind == FK_TWO && r2 != r1) ? 2 : 1;
      std::map<int, int> net;
      for (int k = 0; k < hn.n_reac[i]; ++k) net[hn.reac[3 * i + k] - 1] -= 1;
      for (int k = 0; k < hn.n_prod[i]; ++k) net[hn.prod[4 * i + k] - 1] += 1;
      for (int q = 0; q < ncol; ++q) {
        int j = q == 0 ? r1 : r2;
        for (auto& kv : net) {
          if (kv.second == 0) continue;
          int s = store_index(hn.iperm[kv.first], hn.iperm[j]);
          if (s < 0) { hn.error = "internal: Jacobian entry outside the symbolic pattern"; return false; }
          tgt[s].push_back({i * 2 + q, kv.second});
        }
      }
    }
    std::vector<std::vector<std::pair<int, int>>> rows;
    std::vector<int> row_store;
    for (auto& kv : tgt) { row_store.push_back(kv.first); rows.push_back(kv.second); }
    build_gather(hn.jac, rows, 32);
    // translate row ids to store indices
    for (auto& t : hn.jac.sub_target) if (t >= 0) t = row_store[t];
    for (auto& r : hn.jac.comb_row) r = row_store[r];
  }
  // CSC slot -> storage index (species block only; T row/col are identically zero for evolT=F)
  hn.csc_to_store.assign(hn.NNZ, -1);
  hn.rx_slots.assign((size_t)12 * R, -1);
  {
    std::vector<int> slot_of((size_t)NEQ * NEQ, -1);
    for (int c = 0; c < NEQ; ++c)
      for (int k = hn.ia[c] - 1; k < hn.ia[c + 1] - 1; ++k) {
        int r = hn.ja[k] - 1;
        slot_of[(size_t)c * NEQ + r] = k;
        if (r < n && c < n) hn.csc_to_store[k] = store_index(hn.iperm[r], hn.iperm[c]);
      }
    for (int i = 0; i < R; ++i) {
      const int* sp = &hn.rx_species[(size_t)6 * i];
      for (int q = 0; q < 2; ++q) {
        if (sp[q] < 0) continue;
        for (int k = 0; k < 6; ++k)
          if (sp[k] >= 0) hn.rx_slots[(size_t)12 * i + 6 * q + k] = slot_of[(size_t)sp[q] * NEQ + sp[k]];
      }
    }
  }
  // ---- stand-alone K3 schedule: columns grouped so that a group's partial
  // derivatives fit the shared-memory buffer; hub columns are cut into chunks that
  // accumulate into pd.
  {
    const int CAP = 192;
    HostNet::JacCols& jc = hn.jc;
    std::vector<std::vector<uint32_t>> col_pairs(NEQ);
    for (int i = 0; i < R; ++i) {
      uint32_t w = hn.fw[i];
      int kind = (w >> 20) & 3;
      if (kind == FK_SKIP) continue;
      int r1 = w & 1023, r2 = (w >> 10) & 1023;
      col_pairs[r1].push_back((uint32_t)i | (0u << 16));
      if (kind == FK_TWO && r2 != r1) col_pairs[r2].push_back((uint32_t)i | (1u << 16));
    }
    std::vector<char> written(hn.NNZ, 0);
    jc.grp_pair_ptr.assign(1, 0); jc.grp_slot_ptr.assign(1, 0); jc.slot_ent_ptr.assign(1, 0);
    auto net_coefs = [&](int i) {
      std::map<int, int> net;
      for (int k = 0; k < hn.n_reac[i]; ++k) net[hn.reac[3 * i + k] - 1] -= 1;
      for (int k = 0; k < hn.n_prod[i]; ++k) net[hn.prod[4 * i + k] - 1] += 1;
      return net;
    };
    std::vector<int> slot_of_row(NEQ, -1);
    // emit one group covering pairs [pb,pe) of column list `cols` (single column when chunked)
    auto emit_group = [&](const std::vector<std::pair<int, std::pair<int, int>>>& colranges, int accum) {
      // colranges: (column j, (first pair, last pair))
      int local = 0;
      for (auto& cr : colranges) {
        int j = cr.first;
        for (int k = hn.ia[j] - 1; k < hn.ia[j + 1] - 1; ++k) slot_of_row[hn.ja[k] - 1] = k;
        std::map<int, std::vector<uint32_t>> per_slot;
        for (int p = cr.second.first; p < cr.second.second; ++p) {
          uint32_t pr = col_pairs[j][p];
          jc.pair.push_back(pr);
          auto net = net_coefs(pr & 0xffff);
          for (auto& kv : net) {
            if (kv.second == 0) continue;
            per_slot[slot_of_row[kv.first]].push_back((uint32_t)local | ((uint32_t)(kv.second + 4) << 24));
          }
          ++local;
        }
        for (auto& ps : per_slot) {
          jc.slot_id.push_back(ps.first);
          for (auto e : ps.second) jc.ent.push_back(e);
          jc.slot_ent_ptr.push_back((int)jc.ent.size());
          if (!accum) written[ps.first] = 1;
        }
      }
      jc.grp_pair_ptr.push_back((int)jc.pair.size());
      jc.grp_slot_ptr.push_back((int)jc.slot_id.size());
      jc.grp_accum.push_back(accum);
      jc.max_pairs = std::max(jc.max_pairs, local);
    };
    std::vector<std::pair<int, std::pair<int, int>>> cur; int curp = 0;
    for (int j = 0; j < NEQ; ++j) {
      int np = (int)col_pairs[j].size();
      if (np == 0) continue;
      if (np > CAP) {
        if (!cur.empty()) { emit_group(cur, 0); cur.clear(); curp = 0; }
        // chunk 0 stores, later chunks accumulate; slots first touched by a later chunk are pre-zeroed
        for (int pb = 0, ch = 0; pb < np; pb += CAP, ++ch)
          emit_group({{j, {pb, std::min(np, pb + CAP)}}}, ch > 0 ? 1 : 0);
        continue;
      }
      if (curp + np > CAP) { emit_group(cur, 0); cur.clear(); curp = 0; }
      cur.push_back({j, {0, np}}); curp += np;
    }
    if (!cur.empty()) emit_group(cur, 0);
    jc.ngroups = (int)jc.grp_accum.size();
    for (int k = 0; k < hn.NNZ; ++k) if (!written[k]) jc.zero_slots.push_back(k);
  }
  return true;
}

}  // namespace racg
