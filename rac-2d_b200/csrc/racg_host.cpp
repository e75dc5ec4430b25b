// racg_host.cpp -- host-side network setup of libracg (see racg_host.hpp).
#include "racg_host.hpp"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <map>
#include <numeric>
#include <set>

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
                    const int* dupli_ptr, const int* dupli_list, const racg_cfg* cfg, int nthreads) {
  if (R <= 0 || N <= 0 || N > 1000) { hn.error = "bad R/N (N must be <= 1000)"; return false; }
  if (nthreads < 64 || (nthreads & (nthreads - 1))) { hn.error = "integrator thread count must be a power of two >= 64"; return false; }
  hn.nthreads = nthreads;
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
    // chem_params%R_H2_form_rate_coeff: assigned in the itype 0 branch and in the itype 63 branch
    // when the reactant is gH (src/chemistry.f90:804, 891); the last assignment wins
    if (hn.itype[i] == 0 || (hn.itype[i] == 63 && hn.names[r1] == "gH")) hn.h2form_reac = i;
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
  // ---- streaming K2 schedule (HostNet::RhsChunks)
  {
    HostNet::RhsChunks& rc = hn.rhsc;
    rc.RC = 384; rc.nchunk = (R + rc.RC - 1) / rc.RC; rc.nwarp = 32; rc.spw = (N + rc.nwarp - 1) / rc.nwarp;
    // species -> (warp, slot).  The warps meet at a barrier after every chunk, so what counts is the
    // sum over chunks of the slowest warp's work in that chunk (the network file groups reactions by
    // type: a hub species has most of its terms in a few chunks).  Greedy: heaviest species first,
    // onto the warp with a free slot that raises that sum least (ties: the lighter warp).  Against
    // balancing the totals this shortens the per-chunk maxima from 1.96 x to 1.40 x the mean (A).
    std::vector<int> order(N);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return sp_rows[a].size() > sp_rows[b].size(); });
    std::vector<std::vector<double>> wcost(rc.nwarp, std::vector<double>(rc.nchunk, 0.0));
    std::vector<double> cmax(rc.nchunk, 0.0), sc(rc.nchunk);
    std::vector<int> used(rc.nwarp, 0);
    rc.slot_species.assign((size_t)rc.nwarp * rc.spw, -1);
    for (int sp : order) {
      std::fill(sc.begin(), sc.end(), 0.0);
      for (auto& e : sp_rows[sp]) sc[e.first / rc.RC] += std::abs(e.second);
      for (auto& v : sc) if (v > 0.0) v += 12.0;           // fixed cost of a run
      int best = -1; double bestv = 0.0;
      for (int w = 0; w < rc.nwarp; ++w) {
        if (used[w] >= rc.spw) continue;
        double v = 0.0, tot = 0.0;
        for (int c = 0; c < rc.nchunk; ++c) { const double t = wcost[w][c] + sc[c]; v += std::max(cmax[c], t); tot += t; }
        v += 1e-3 * tot;
        if (best < 0 || v < bestv) { best = w; bestv = v; }
      }
      rc.slot_species[(size_t)best * rc.spw + used[best]++] = sp;
      for (int c = 0; c < rc.nchunk; ++c) { wcost[best][c] += sc[c]; cmax[c] = std::max(cmax[c], wcost[best][c]); }
    }
    // run lists: entries are byte offsets of the chunk's rows (row pitch 256 B), 4 per 16-byte group
    rc.off.assign((size_t)rc.nwarp * rc.nchunk, 0); rc.nrun.assign((size_t)rc.nwarp * rc.nchunk, 0);
    rc.len4.assign((size_t)rc.nwarp * rc.nchunk, 0);
    for (int w = 0; w < rc.nwarp; ++w)
      for (int c = 0; c < rc.nchunk; ++c) {
        while (rc.stream.size() % 4) rc.stream.push_back(0u);
        rc.off[(size_t)w * rc.nchunk + c] = (uint32_t)rc.stream.size();
        int nrun = 0;
        std::vector<uint32_t> hdr, ent;
        for (int k = 0; k < rc.spw; ++k) {
          const int sp = rc.slot_species[(size_t)w * rc.spw + k];
          if (sp < 0) continue;
          std::vector<uint32_t> em, ep;
          for (auto& e : sp_rows[sp]) {
            if (e.first / rc.RC != c) continue;
            const uint32_t loc = (uint32_t)(e.first % rc.RC) * 256u;
            for (int m = 0; m < std::abs(e.second); ++m) (e.second < 0 ? em : ep).push_back(loc);
          }
          if (em.empty() && ep.empty()) continue;
          while (em.size() % 4) em.push_back((uint32_t)rc.RC * 256u);
          while (ep.size() % 4) ep.push_back((uint32_t)rc.RC * 256u);
          hdr.push_back((uint32_t)k | ((uint32_t)(em.size() / 4) << 5) | ((uint32_t)(ep.size() / 4) << 18));
          ent.insert(ent.end(), em.begin(), em.end());
          ent.insert(ent.end(), ep.begin(), ep.end());
          ++nrun;
        }
        // layout of one (warp, chunk) list: headers padded to a multiple of 4 words, then the entries
        while (hdr.size() % 4) hdr.push_back(0u);
        rc.stream.insert(rc.stream.end(), hdr.begin(), hdr.end());
        rc.stream.insert(rc.stream.end(), ent.begin(), ent.end());
        rc.nrun[(size_t)w * rc.nchunk + c] = nrun;
        const int l4 = (int)((hdr.size() + ent.size()) / 4);
        rc.len4[(size_t)w * rc.nchunk + c] = l4;
        rc.max_len4 = std::max(rc.max_len4, l4);
      }
    for (int q = 0; q < 8; ++q) rc.stream.push_back(0u);
    // flux lists per chunk, sorted by kind: word = local row | r1 << 9 | (r2 or saturation index) << 19
    rc.fl_off.assign((size_t)rc.nchunk * 4, 0);
    for (int c = 0; c < rc.nchunk; ++c) {
      for (int kind = 0; kind < 3; ++kind) {
        rc.fl_off[(size_t)c * 4 + kind] = (int)rc.flux.size();
        for (int i = c * rc.RC; i < std::min(R, (c + 1) * rc.RC); ++i) {
          const uint32_t w = hn.fw[i];
          if ((int)((w >> 20) & 3) != kind) continue;
          const uint32_t second = (kind == FK_SAT) ? (w >> 22) : ((w >> 10) & 1023u);
          rc.flux.push_back((uint32_t)(i - c * rc.RC) | ((w & 1023u) << 9) | (second << 19));
        }
      }
      rc.fl_off[(size_t)c * 4 + 3] = (int)rc.flux.size();
    }
    rc.flux.push_back(0u);
  }

  // ---- species-block pattern and fill-reducing ordering
  const int n = N;
  const int W = (n + 63) / 64;
  std::vector<uint64_t> A((size_t)n * W, 0);   // A[row][col] bits, unsymmetric pattern incl. diagonal
  auto setb = [&](std::vector<uint64_t>& M, int r, int c) { M[(size_t)r * W + (c >> 6)] |= (1ull << (c & 63)); };
  auto getb = [&](const std::vector<uint64_t>& M, int r, int c) { return (M[(size_t)r * W + (c >> 6)] >> (c & 63)) & 1ull; };
  for (int c = 0; c < n; ++c)
    for (int k = hn.ia[c] - 1; k < hn.ia[c + 1] - 1; ++k) { int r = hn.ja[k] - 1; if (r < n) setb(A, r, c); }
  for (int i = 0; i < n; ++i) setb(A, i, i);
  std::vector<uint64_t> G0((size_t)n * W, 0);   // symmetric graph of M + M^T (no self loops)
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) if (r != c && (getb(A, r, c) || getb(A, c, r))) { setb(G0, r, c); setb(G0, c, r); }
  // elimination on the bitset graph; multi = ratio for multiple elimination of an
  // independent set of near-minimum-degree vertices per round (shallower elimination
  // tree = fewer solve levels); nstop = number of pivots to take this way
  auto eliminate = [&](std::vector<uint64_t>& G, std::vector<char>& done, std::vector<int>& order, int v) {
    done[v] = 1; order.push_back(v);
    std::vector<uint64_t> nb(G.begin() + (size_t)v * W, G.begin() + (size_t)v * W + W);
    for (int a = 0; a < n; ++a) {
      if (!((nb[a >> 6] >> (a & 63)) & 1ull)) continue;
      uint64_t* ga = &G[(size_t)a * W];
      for (int w = 0; w < W; ++w) ga[w] |= nb[w];
      ga[a >> 6] &= ~(1ull << (a & 63));
      ga[v >> 6] &= ~(1ull << (v & 63));
    }
    for (int w = 0; w < W; ++w) G[(size_t)v * W + w] = 0;
  };
  auto degree = [&](const std::vector<uint64_t>& G, int v) { int d = 0; for (int w = 0; w < W; ++w) d += __builtin_popcountll(G[(size_t)v * W + w]); return d; };
  auto min_degree_order = [&](double multi, int nmulti, int* clique_start) {
    std::vector<uint64_t> G = G0;
    std::vector<char> done(n, 0);
    std::vector<int> order;
    if (clique_start) *clique_start = -1;
    while ((int)order.size() < n) {
      int dmin = n + 1;
      for (int v = 0; v < n; ++v) if (!done[v]) dmin = std::min(dmin, degree(G, v));
      if (clique_start && *clique_start < 0 && dmin >= n - (int)order.size() - 1) *clique_start = (int)order.size();
      if ((int)order.size() < nmulti && multi > 1.0) {
        std::vector<std::pair<int, int>> cand;
        for (int v = 0; v < n; ++v) if (!done[v]) { int d = degree(G, v); if (d <= multi * dmin) cand.push_back({d, v}); }
        std::sort(cand.begin(), cand.end());
        std::vector<uint64_t> blocked(W, 0);
        std::vector<int> chosen;
        for (auto& dv : cand) {
          int v = dv.second;
          if ((blocked[v >> 6] >> (v & 63)) & 1ull) continue;
          if ((int)order.size() + (int)chosen.size() >= nmulti) break;
          chosen.push_back(v);
          for (int w = 0; w < W; ++w) blocked[w] |= G[(size_t)v * W + w];
          blocked[v >> 6] |= 1ull << (v & 63);
        }
        for (int v : chosen) eliminate(G, done, order, v);
      } else {
        int best = -1, bd = n + 1;
        for (int v = 0; v < n; ++v) if (!done[v]) { int d = degree(G, v); if (d < bd) { bd = d; best = v; } }
        eliminate(G, done, order, best);
      }
    }
    return order;
  };
  // symbolic LU of the permuted pattern for a given order -> F (bitset rows)
  auto symbolic = [&](const std::vector<int>& perm, std::vector<uint64_t>& F) {
    F.assign((size_t)n * W, 0);
    for (int r = 0; r < n; ++r)
      for (int c = 0; c < n; ++c) if (getb(A, perm[r], perm[c])) setb(F, r, c);
    for (int i = 0; i < n; ++i) {
      uint64_t* fi = &F[(size_t)i * W];
      for (int k = 0; k < i; ++k) {
        if (!((fi[k >> 6] >> (k & 63)) & 1ull)) continue;
        const uint64_t* fk = &F[(size_t)k * W];
        for (int w = (k >> 6); w < W; ++w) {
          uint64_t m = fk[w];
          if (w == (k >> 6)) m &= ~((2ull << (k & 63)) - 1ull);
          fi[w] |= m;
        }
      }
    }
  };
  int clique_start = -1;
  std::vector<int> order0 = min_degree_order(1.0, 0, &clique_start);
  if (clique_start < 0) clique_start = n - 1;
  int nt0 = ((n - clique_start + 15) / 16) * 16;
  nt0 = std::max(16, std::min(std::min(nt0, 128), (n / 16) * 16));
  const double multi_ratio = 1.0;   // ratio of the multiple-elimination rounds (1 = plain minimum degree)
  std::vector<uint64_t> F;
  const size_t smem_budget = 227 * 1024 - 2048;
  bool placed = false;
  for (int nt = nt0; nt >= 16 && !placed; nt -= 16) {
    hn.nt = nt; hn.nh = n - nt;
    hn.perm = min_degree_order(multi_ratio, hn.nh, nullptr);
    symbolic(hn.perm, F);
    int n_hh = 0;
    for (int i = 0; i < hn.nh; ++i) for (int c = 0; c < hn.nh; ++c) if (getb(F, i, c) || c == i) ++n_hh;
    // minimum shared memory of the integrator (racg_integrate.cu make_layout): 2 length-n
    // vectors + 1/pivots + tail (ld = nt+1) + scratch; the head x head block and the
    // factorisation's U_B copy go to the L2 workspace when they do not fit as well
    (void)n_hh;
    size_t scratch = std::max<size_t>((size_t)R + 2048, (size_t)8 * n);
    size_t need = 8 * ((size_t)2 * n + hn.nh + 64 + (size_t)(nt + 1) * nt + scratch);
    if (need <= smem_budget) placed = true;
  }
  if (!placed) { hn.error = "network too large for the integrator's shared-memory layout"; return false; }
  const int nh = hn.nh, nt = hn.nt;
  hn.iperm.assign(n, 0);
  for (int i = 0; i < n; ++i) hn.iperm[hn.perm[i]] = i;
  hn.nnz_lu = 0;
  for (int i = 0; i < n; ++i) for (int w = 0; w < W; ++w) hn.nnz_lu += __builtin_popcountll(F[(size_t)i * W + w]);
  // ---- storage layout shared by the Jacobian (J) and factor (LU) value arrays
  // head rows: [L_A | diag | U_A] (hh, resident in shared memory) and U_B (head rows x tail
  // columns, CSR; in shared memory only while the matrix is being factorised)
  hn.hh_ptr.assign(nh + 1, 0); hn.hh_nl.assign(nh, 0); hn.hh_col.clear();
  hn.ub_ptr.assign(nh + 1, 0); hn.ub_col.clear();
  hn.lc_ptr.assign(nt + 1, 0); hn.lc_col.clear();
  for (int i = 0; i < nh; ++i) {
    int nl = 0;
    for (int c = 0; c < nh; ++c) if (getb(F, i, c) || c == i) { hn.hh_col.push_back((uint16_t)c); if (c < i) ++nl; }
    hn.hh_nl[i] = nl; hn.hh_ptr[i + 1] = (int)hn.hh_col.size();
    for (int c = nh; c < n; ++c) if (getb(F, i, c)) hn.ub_col.push_back((uint16_t)(c - nh));
    hn.ub_ptr[i + 1] = (int)hn.ub_col.size();
  }
  for (int a = 0; a < nt; ++a) {
    for (int c = 0; c < nh; ++c) if (getb(F, nh + a, c)) hn.lc_col.push_back((uint16_t)c);
    hn.lc_ptr[a + 1] = (int)hn.lc_col.size();
  }
  hn.n_hh = (int)hn.hh_col.size(); hn.n_ub = (int)hn.ub_col.size(); hn.n_lc = (int)hn.lc_col.size();
  hn.ldt = nt + 1;
  hn.o_ub = hn.n_hh; hn.o_lc = hn.o_ub + hn.n_ub; hn.o_tl = hn.o_lc + hn.n_lc;
  hn.nstore = hn.o_tl + hn.ldt * nt;
  // ELL copies of U_B and L_C for the solve's SpMV passes
  auto build_ell = [&](const std::vector<int>& ptr, const std::vector<uint16_t>& col, int nrows, HostNet::Ell& e,
                       std::vector<int>& ellpos) {
    std::vector<std::vector<std::pair<int, int>>> rows(nrows);
    for (int i = 0; i < nrows; ++i) for (int q = ptr[i]; q < ptr[i + 1]; ++q) rows[i].push_back({q, 0});
    Gather g;
    const int SEG = 16;    // sub-row length: one batch of loads per block in the solves
    build_gather(g, rows, SEG);
    e.nblk = g.nblk; e.npartial = g.npartial; e.ncombine = g.ncombine;
    e.blk_off = g.blk_off; e.blk_width = g.blk_width; e.sub_target = g.sub_target;
    e.comb_row = g.comb_row; e.comb_ptr = g.comb_ptr;
    e.nval = (int)g.ent.size();
    e.col.assign(g.ent.size(), 0);
    ellpos.assign(col.size(), -1);
    // padding entries were created with idx 0 / coef code 4; real entries carry coef code 4 too
    // (coef 0), so mark real ones through a second pass
    std::vector<char> real(g.ent.size(), 0);
    {
      // re-derive positions exactly as build_gather laid them out
      struct Sub { int row, first, len; };
      std::vector<Sub> subs;
      for (int r = 0; r < nrows; ++r) {
        int len = (int)rows[r].size();
        for (int f = 0; f < len; f += SEG) subs.push_back({r, f, std::min(SEG, len - f)});
      }
      std::stable_sort(subs.begin(), subs.end(), [](const Sub& x, const Sub& y) { return x.len > y.len; });
      for (size_t s = 0; s < subs.size(); ++s) {
        int b = (int)(s / 32), l = (int)(s % 32);
        for (int j = 0; j < subs[s].len; ++j) {
          size_t pos = (size_t)g.blk_off[b] + (size_t)j * 32 + l;
          int q = rows[subs[s].row][subs[s].first + j].first;
          e.col[pos] = col[q];
          ellpos[q] = (int)pos;
          real[pos] = 1;
        }
      }
    }
  };
  build_ell(hn.ub_ptr, hn.ub_col, nh, hn.ubE, hn.ub_ellpos);
  build_ell(hn.lc_ptr, hn.lc_col, nt, hn.lcE, hn.lc_ellpos);
  // levels over head rows: forward (L_A dependencies) and backward (U_A dependencies)
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
  auto first_thin = [&](const std::vector<int>& ptr) {
    int nlev = (int)ptr.size() - 1, f = nlev;
    while (f > 0 && ptr[f] - ptr[f - 1] <= 32) --f;
    return f;   // levels [f, nlev) all have <= 32 rows
  };
  {
    std::vector<int> lev(nh, 0);
    for (int i = 0; i < nh; ++i) {
      int l = 0;
      for (int q = hn.hh_ptr[i]; q < hn.hh_ptr[i] + hn.hh_nl[i]; ++q) l = std::max(l, lev[hn.hh_col[q]] + 1);
      lev[i] = l;
    }
    make_levels(lev, nh, hn.flev_ptr, hn.flev_rows);
    hn.head_level = lev;
    hn.nfat_f = first_thin(hn.flev_ptr);
    std::vector<int> levu(nh, 0);
    for (int i = nh - 1; i >= 0; --i) {
      int l = 0;
      for (int q = hn.hh_ptr[i] + hn.hh_nl[i] + 1; q < hn.hh_ptr[i + 1]; ++q) l = std::max(l, levu[hn.hh_col[q]] + 1);
      levu[i] = l;
    }
    make_levels(levu, nh, hn.su_ptr, hn.su_rows);
    hn.nfat_b = first_thin(hn.su_ptr);
  }
  hn.pivmeta.assign((size_t)4 * nh, 0); hn.fmeta.assign((size_t)4 * nh, 0); hn.bmeta.assign((size_t)4 * nh, 0);
  for (int k = 0; k < nh; ++k) {
    hn.pivmeta[4 * k + 0] = hn.hh_ptr[k] + hn.hh_nl[k] + 1;                       // start of U_A(k,:) in hh
    hn.pivmeta[4 * k + 1] = hn.hh_ptr[k + 1] - (hn.hh_ptr[k] + hn.hh_nl[k] + 1);    // its length
    hn.pivmeta[4 * k + 2] = hn.ub_ptr[k];                                           // start of U_B(k,:)
    hn.pivmeta[4 * k + 3] = hn.ub_ptr[k + 1] - hn.ub_ptr[k];                        // its length
  }
  for (int p = 0; p < nh; ++p) {
    int i = hn.flev_rows[p];
    hn.fmeta[4 * p + 0] = i; hn.fmeta[4 * p + 1] = hn.hh_ptr[i]; hn.fmeta[4 * p + 2] = hn.hh_nl[i];
    i = hn.su_rows[p];
    hn.bmeta[4 * p + 0] = i; hn.bmeta[4 * p + 1] = hn.hh_ptr[i] + hn.hh_nl[i] + 1;
    hn.bmeta[4 * p + 2] = hn.hh_ptr[i + 1] - (hn.hh_ptr[i] + hn.hh_nl[i] + 1);
  }
  // tail rows in decreasing L_C length (longest first) for the Schur phase
  hn.tail_order.resize(nt);
  std::iota(hn.tail_order.begin(), hn.tail_order.end(), 0);
  std::stable_sort(hn.tail_order.begin(), hn.tail_order.end(), [&](int x, int y) {
    return hn.lc_ptr[x + 1] - hn.lc_ptr[x] > hn.lc_ptr[y + 1] - hn.lc_ptr[y]; });

  // ---- Jacobian gather: J(i,j) = sum_r coef_i(r) * dflux[r][which(j)]  (src/disk.f90:4765-4875)
  auto store_index = [&](int pi, int pj) -> int {
    if (pi >= nh && pj >= nh) return hn.o_tl + (pj - nh) * hn.ldt + (pi - nh);
    if (pi < nh && pj < nh) {
      auto lo = hn.hh_col.begin() + hn.hh_ptr[pi], hi = hn.hh_col.begin() + hn.hh_ptr[pi + 1];
      auto it = std::lower_bound(lo, hi, (uint16_t)pj);
      return (it == hi || *it != pj) ? -1 : (int)(it - hn.hh_col.begin());
    }
    if (pi < nh) {
      auto lo = hn.ub_col.begin() + hn.ub_ptr[pi], hi = hn.ub_col.begin() + hn.ub_ptr[pi + 1];
      auto it = std::lower_bound(lo, hi, (uint16_t)(pj - nh));
      return (it == hi || *it != pj - nh) ? -1 : hn.o_ub + (int)(it - hn.ub_col.begin());
    }
    auto lo = hn.lc_col.begin() + hn.lc_ptr[pi - nh], hi = hn.lc_col.begin() + hn.lc_ptr[pi - nh + 1];
    auto it = std::lower_bound(lo, hi, (uint16_t)pj);
    return (it == hi || *it != pj) ? -1 : hn.o_lc + (int)(it - hn.lc_col.begin());
  };
  {
    // two passes so that one R-sized scratch suffices: pass q uses d(flux)/d(y_rq)
    std::map<int, std::vector<std::pair<int, int>>> tgt[2];   // store idx -> (r, coef)
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
          tgt[q][s].push_back({i, kv.second});
        }
      }
    }
    for (int q = 0; q < 2; ++q) {
      std::vector<std::vector<std::pair<int, int>>> rows;
      std::vector<int> row_store;
      for (auto& kv : tgt[q]) { row_store.push_back(kv.first); rows.push_back(kv.second); }
      Gather& g = hn.jac[q];
      build_gather(g, rows, 32);
      for (auto& t : g.sub_target) if (t >= 0) t = row_store[t];
      for (auto& r : g.comb_row) r = row_store[r];
      // pass 1 adds to what pass 0 stored; targets untouched by pass 0 are stored
      g.sub_add.assign(g.sub_target.size(), 0);
      g.comb_add.assign(g.comb_row.size(), 0);
      if (q == 1) {
        for (size_t k = 0; k < g.sub_target.size(); ++k)
          if (g.sub_target[k] >= 0 && tgt[0].count(g.sub_target[k])) g.sub_add[k] = 1;
        for (size_t k = 0; k < g.comb_row.size(); ++k) if (tgt[0].count(g.comb_row[k])) g.comb_add[k] = 1;
      }
    }
  }
  // CSC slot -> storage index (species block only; T row/col are identically zero for evolT=F)
  hn.csc_to_store.assign(hn.NNZ, -1);
  for (int c = 0; c < n; ++c)
    for (int k = hn.ia[c] - 1; k < hn.ia[c + 1] - 1; ++k) {
      int r = hn.ja[k] - 1;
      if (r < n) hn.csc_to_store[k] = store_index(hn.iperm[r], hn.iperm[c]);
    }
  // ---- level-parallel factorisation schedule (HostNet::LevelLU) and staged solves
  if (hn.nstore + 2 < 65535) {
    HostNet::LevelLU& g = hn.glu;
    g.zpos = hn.nstore;
    std::vector<int> lev(nh, 0);
    for (int k = 0; k < nh; ++k) {
      int l = 0;
      for (int c = 0; c < k; ++c) if (getb(F, k, c) || getb(F, c, k)) l = std::max(l, lev[c] + 1);
      lev[k] = l;
    }
    g.nlev = 0;
    for (int k = 0; k < nh; ++k) g.nlev = std::max(g.nlev, lev[k] + 1);
    std::vector<std::vector<int>> members(g.nlev);
    for (int k = 0; k < nh; ++k) members[lev[k]].push_back(k);
    bool ok = true;
    for (int L = 0; L < g.nlev && ok; ++L) {
      const bool rank1 = members[L].size() == 1;
      g.lvl.insert(g.lvl.end(), {(int)g.piv.size(), (int)g.mul.size(), (int)g.grp.size() / 4,
                                 rank1 ? (int)g.r1.size() / 8 : -1});
      if (rank1) {
        const int k = members[L][0];
        g.piv.push_back((uint32_t)store_index(k, k) | ((uint32_t)k << 16));
        std::vector<int> rws, cols;
        for (int i = k + 1; i < n; ++i) if (getb(F, i, k)) rws.push_back(i);
        const int ua0 = hn.hh_ptr[k] + hn.hh_nl[k] + 1, nua = hn.hh_ptr[k + 1] - ua0;
        const int ub0 = hn.o_ub + hn.ub_ptr[k], nub = hn.ub_ptr[k + 1] - hn.ub_ptr[k];
        const int nua4 = (nua + 3) / 4, nub4 = (nub + 3) / 4, nj4 = nua4 + nub4;
        // column of chunk j4, slot c (permuted index) or -1 for padding
        auto col_of = [&](int j4, int c) -> int {
          if (j4 < nua4) { int q = 4 * j4 + c; return q < nua ? (int)hn.hh_col[ua0 + q] : -1; }
          int q = 4 * (j4 - nua4) + c; return q < nub ? nh + (int)hn.ub_col[hn.ub_ptr[k] + q] : -1;
        };
        for (int i : rws) g.mul.push_back((uint32_t)store_index(i, k) | ((uint32_t)store_index(k, k) << 16));
        const int nr = (int)rws.size(), nrc = (nr + 31) / 32;
        g.r1.insert(g.r1.end(), {ua0, nua4, ub0, nub4, (int)(g.r1tgt.size() / 4), nr, nj4, 0});
        for (int rc = 0; rc < nrc; ++rc)
          for (int j4 = 0; j4 < nj4; ++j4)
            for (int l = 0; l < 32; ++l)
              for (int c = 0; c < 4; ++c) {
                const int m = rc * 32 + l, j = col_of(j4, c);
                int t = g.zpos + 1;
                if (m < nr && j >= 0) { t = store_index(rws[m], j); if (t < 0) ok = false; ++g.npairs; }
                g.r1tgt.push_back((uint16_t)t);
              }
        continue;
      }
      std::map<int, std::vector<uint32_t>> upd;     // target position -> pairs
      for (int k : members[L]) {
        g.piv.push_back((uint32_t)store_index(k, k) | ((uint32_t)k << 16));
        std::vector<int> cols;
        for (int j = k + 1; j < n; ++j) if (getb(F, k, j)) cols.push_back(j);
        for (int i = k + 1; i < n; ++i) {
          if (!getb(F, i, k)) continue;
          const int pl = store_index(i, k);
          g.mul.push_back((uint32_t)pl | ((uint32_t)store_index(k, k) << 16));
          for (int j : cols) {
            const int pt = store_index(i, j), pu = store_index(k, j);
            if (pt < 0 || pu < 0 || pl < 0) { ok = false; break; }
            upd[pt].push_back((uint32_t)pl | ((uint32_t)pu << 16));
          }
        }
      }
      std::vector<std::pair<int, int>> order;   // (-count, target)
      for (auto& kv : upd) { order.push_back({-(int)kv.second.size(), kv.first}); g.npairs += (long)kv.second.size(); }
      std::sort(order.begin(), order.end());
      // blocks of 32 targets; consecutive blocks of equal width form a group
      for (size_t s0 = 0; s0 < order.size(); s0 += 32) {
        const int width = -order[s0].first;
        const size_t ng = g.grp.size();
        if (ng == (size_t)4 * g.lvl[g.lvl.size() - 2] || g.grp[ng - 4] != width)
          g.grp.insert(g.grp.end(), {width, 0, (int)g.ent.size(), (int)g.tgt.size()});
        g.grp[g.grp.size() - 3] += 1;
        const size_t off = g.ent.size();
        g.ent.resize(off + (size_t)width * 32, (uint32_t)g.zpos | ((uint32_t)g.zpos << 16));
        for (int l = 0; l < 32; ++l) {
          if (s0 + l >= order.size()) { g.tgt.push_back(0xFFFF); continue; }
          const int t = order[s0 + l].second;
          g.tgt.push_back((uint16_t)t);
          const auto& pr = upd[t];
          for (size_t j = 0; j < pr.size(); ++j) g.ent[off + j * 32 + l] = pr[j];
        }
      }
    }
    g.lvl.insert(g.lvl.end(), {(int)g.piv.size(), (int)g.mul.size(), (int)g.grp.size() / 4, 0});
    if (!ok) g = HostNet::LevelLU();
    if (ok) {
      HostNet::SolveSched& ss = hn.ss;
      // S: the longest suffix of levels holding <= 96 rows
      int thr = g.nlev, cnt = 0;
      while (thr > 0 && cnt + (int)members[thr - 1].size() <= 96) { cnt += (int)members[thr - 1].size(); --thr; }
      std::vector<int> Srows, blkOf(nh, -1), locOf(nh, 0);
      for (int k = 0; k < nh; ++k) if (lev[k] >= thr) Srows.push_back(k);
      ss.nblkS = ((int)Srows.size() + 31) / 32;
      for (size_t q = 0; q < Srows.size(); ++q) { blkOf[Srows[q]] = (int)(q / 32); locOf[Srows[q]] = (int)(q % 32); }
      std::vector<uint32_t> t_ent; std::vector<uint16_t> t_rp, t_rows;
      auto add_stage = [&](int kind, const std::vector<int>& rows, const std::vector<std::vector<uint32_t>>& ents, int inv) {
        const int nrows = (int)rows.size();
        int maxlen = 0;
        for (auto& e : ents) maxlen = std::max(maxlen, (int)e.size());
        int lg;
        int lgblk = 0;
        while ((32 << (lgblk + 1)) <= hn.nthreads) ++lgblk;      // lanes per row of a 32-row block stage
        if (kind == 2) lg = lgblk;
        else { lg = 5; while (lg > 0 && (nrows << lg) > hn.nthreads) --lg; while (lg > 0 && (1 << (lg - 1)) >= std::max(1, maxlen)) --lg; }
        const int row_off = (int)t_rows.size(), rp_off = (int)t_rp.size();
        for (int r = 0; r < nrows; ++r) {
          t_rows.push_back((uint16_t)rows[r]);
          t_rp.push_back((uint16_t)t_ent.size());
          for (uint32_t e : ents[r]) t_ent.push_back(e);
        }
        t_rp.push_back((uint16_t)t_ent.size());
        ss.st.insert(ss.st.end(), {kind | (lg << 8), nrows | (inv << 16), row_off, rp_off});
      };
      auto row_entries = [&](int i, bool upper, int skip_blk) {
        std::vector<uint32_t> e;
        const int b0 = upper ? hn.hh_ptr[i] + hn.hh_nl[i] + 1 : hn.hh_ptr[i];
        const int b1 = upper ? hn.hh_ptr[i + 1] : hn.hh_ptr[i] + hn.hh_nl[i];
        for (int q = b0; q < b1; ++q) {
          const int c = hn.hh_col[q];
          if (skip_blk >= 0 && blkOf[c] == skip_blk) continue;
          e.push_back((uint32_t)q | ((uint32_t)c << 16));
        }
        return e;
      };
      auto add_level = [&](int kind, const std::vector<int>& rows, bool upper) {
        const size_t NTH = (size_t)hn.nthreads;
        for (size_t r0 = 0; r0 < rows.size(); r0 += NTH) {
          std::vector<int> rr(rows.begin() + r0, rows.begin() + std::min(rows.size(), r0 + NTH));
          std::vector<std::vector<uint32_t>> ee;
          for (int i : rr) ee.push_back(row_entries(i, upper, -1));
          add_stage(kind, rr, ee, 0);
        }
      };
      auto add_block = [&](int b, bool upper) {
        std::vector<int> rr(Srows.begin() + 32 * b, Srows.begin() + std::min<size_t>(Srows.size(), 32 * b + 32));
        std::vector<std::vector<uint32_t>> ee;
        for (int i : rr) ee.push_back(row_entries(i, upper, b));
        add_stage(2, rr, ee, b);
      };
      // forward: levels 1..thr-1 (level 0 has nothing to subtract), then the S blocks
      for (int L = 1; L < thr; ++L) add_level(0, members[L], false);
      for (int b = 0; b < ss.nblkS; ++b) add_block(b, false);
      ss.nf = (int)ss.st.size() / 4;
      // backward: S blocks in reverse, then the other rows by their depth below S
      for (int b = ss.nblkS - 1; b >= 0; --b) add_block(b, true);
      {
        std::vector<int> bl(nh, 0);
        int nbl = 0;
        for (int i = nh - 1; i >= 0; --i) {
          if (blkOf[i] >= 0) continue;
          int l = 0;
          for (int q = hn.hh_ptr[i] + hn.hh_nl[i] + 1; q < hn.hh_ptr[i + 1]; ++q) {
            const int c = hn.hh_col[q];
            if (blkOf[c] < 0) l = std::max(l, bl[c] + 1);
          }
          bl[i] = l; nbl = std::max(nbl, l + 1);
        }
        std::vector<std::vector<int>> bm(nbl);
        for (int i = 0; i < nh; ++i) if (blkOf[i] < 0) bm[bl[i]].push_back(i);
        for (int L = 0; L < nbl; ++L) add_level(1, bm[L], true);
      }
      ss.nb = (int)ss.st.size() / 4 - ss.nf;
      // dense copies of the S diagonal blocks: unit-lower part -> tile 2b, upper part -> tile 2b+1
      for (int i : Srows)
        for (int q = hn.hh_ptr[i]; q < hn.hh_ptr[i + 1]; ++q) {
          const int c = hn.hh_col[q];
          if (blkOf[c] != blkOf[i]) continue;
          ss.ext.push_back((uint32_t)q | ((uint32_t)(blkOf[i] * 33 * 32 + locOf[c] * 33 + locOf[i]) << 16));
        }
      // one blob the kernel stages into shared memory: entries | row pointers | row ids
      if (t_ent.size() >= 65535) { hn.glu = HostNet::LevelLU(); ss = HostNet::SolveSched(); }
      else {
        ss.nent = (int)t_ent.size(); ss.nrp = (int)t_rp.size(); ss.nrows = (int)t_rows.size();
        ss.blob = t_ent;
        auto pack16 = [&](const std::vector<uint16_t>& v) {
          for (size_t q = 0; q < v.size(); q += 2)
            ss.blob.push_back((uint32_t)v[q] | ((q + 1 < v.size() ? (uint32_t)v[q + 1] : 0u) << 16));
        };
        pack16(t_rp); pack16(t_rows);
      }
    }
  }
  // ---- stand-alone K3 schedule: columns grouped so that a group's partial
  // derivatives fit the shared-memory buffer; hub columns are cut into chunks that
  // accumulate into pd.
  {
    const int CAP = 192;     // pairs per group: 192 rows x 512 B (64-cell tiles) fit twice in an SM's shared memory
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
      const size_t pair0 = jc.pair.size(), ent0 = jc.ent.size();
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
      // pairs of a group ordered by flux kind (one-body | two-body | saturating): branch-free loops
      // in jac_kernel_pipe, nothing to do for the one-body pairs (the derivative is the rate)
      {
        auto rank = [&](uint32_t pr) { const int kind = (hn.fw[pr & 0xffff] >> 20) & 3; return kind == FK_ONE ? 0 : kind == FK_TWO ? 1 : 2; };
        std::vector<int> perm(local), inv(local);
        std::iota(perm.begin(), perm.end(), 0);
        std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return rank(jc.pair[pair0 + a]) < rank(jc.pair[pair0 + b]); });
        std::vector<uint32_t> old(jc.pair.begin() + pair0, jc.pair.end());
        int n1 = 0, n2 = 0;
        for (int q = 0; q < local; ++q) {
          jc.pair[pair0 + q] = old[perm[q]]; inv[perm[q]] = q;
          const int rk = rank(old[perm[q]]);
          n1 += rk == 0; n2 += rk <= 1;
        }
        for (size_t e = ent0; e < jc.ent.size(); ++e) jc.ent[e] = (jc.ent[e] & 0xff000000u) | (uint32_t)inv[jc.ent[e] & 0xffffffu];
        jc.grp_two_ptr.push_back((int)pair0 + n1);
        jc.grp_sat_ptr.push_back((int)pair0 + n2);
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
    // repack for jac_kernel_pipe
    // pair word x: two-body -> row of the other reactant | row of this column's reactant << 10 | (r1 == r2) << 20;
    //              saturating -> r1 | saturation constant index << 10;  y: reaction (rate row)
    for (uint32_t pr : jc.pair) {
      const uint32_t w = hn.fw[pr & 0xffff];
      const int kind = (w >> 20) & 3, which = pr >> 16;
      const uint32_t r1 = w & 1023, r2 = (w >> 10) & 1023;
      uint32_t x = 0;
      if (kind == FK_TWO) x = (which == 0 ? r2 : r1) | ((which == 0 ? r1 : r2) << 10) | ((r1 == r2 ? 1u : 0u) << 20);
      else if (kind == FK_SAT) x = r1 | ((w >> 22) << 10);
      jc.pairw.push_back(x); jc.pairw.push_back(pr & 0xffff);
    }
    // entry of the repacked lists: local pair index | high 16 bits of the coefficient as a double
    // (byte offset of the pair's 512-byte row in the 64-cell derivative buffer, 17 bits | top 15 bits of the double)
    auto cf15 = [](int c) { const double d = (double)c; uint64_t b; memcpy(&b, &d, 8); return (uint32_t)(b >> 49) << 17; };
    auto ent_of = [&](uint32_t local, int c) { return (local * 512u) | cf15(c); };
    const uint32_t padent = ent_of((uint32_t)jc.max_pairs, 1);
    for (int g = 0; g < jc.ngroups; ++g) {
      std::vector<std::pair<int, int>> ord;   // (-entries, listed slot)
      for (int s2 = jc.grp_slot_ptr[g]; s2 < jc.grp_slot_ptr[g + 1]; ++s2)
        ord.push_back({-(jc.slot_ent_ptr[s2 + 1] - jc.slot_ent_ptr[s2]), s2});
      std::sort(ord.begin(), ord.end());
      for (auto& o : ord) {
        const int s2 = o.second, e0 = jc.slot_ent_ptr[s2], e1 = jc.slot_ent_ptr[s2 + 1];
        const int n4 = (e1 - e0 + 3) / 4;
        jc.slotw.push_back((uint32_t)jc.slot_id[s2]);
        jc.slotw.push_back((uint32_t)(jc.ent4.size() / 4) | ((uint32_t)n4 << 24));
        for (int e = e0; e < e0 + 4 * n4; ++e)
          jc.ent4.push_back(e < e1 ? ent_of(jc.ent[e] & 0xffffffu, (int)(jc.ent[e] >> 24) - 4) : padent);
      }
    }
  }
  return true;
}

std::string describe_host_net(const HostNet& hn) {
  char buf[4096];
  int o = snprintf(buf, sizeof(buf),
                   "n=%d nh=%d nt=%d n_hh=%d n_ub=%d n_lc=%d nstore=%d nnz_lu=%d flev=%d su=%d\n"
                   "level LU: nlev=%d pairs=%ld ent=%zu mul=%zu groups=%zu rank1=%zu (target slots %zu)\n"
                   "solve stages fwd=%d bwd=%d blocksS=%d ent=%d blob=%zu words; ELL blocks U_B %d L_C %d\nstages:",
                   hn.n, hn.nh, hn.nt, hn.n_hh, hn.n_ub, hn.n_lc, hn.nstore, hn.nnz_lu, (int)hn.flev_ptr.size() - 1,
                   (int)hn.su_ptr.size() - 1, hn.glu.nlev, hn.glu.npairs, hn.glu.ent.size(), hn.glu.mul.size(),
                   hn.glu.grp.size() / 4, hn.glu.r1.size() / 8, hn.glu.r1tgt.size(), hn.ss.nf, hn.ss.nb, hn.ss.nblkS,
                   hn.ss.nent, hn.ss.blob.size(), hn.ubE.nblk, hn.lcE.nblk);
  for (size_t q = 0; q < hn.ss.st.size() && o < (int)sizeof(buf) - 64; q += 4)
    o += snprintf(buf + o, sizeof(buf) - o, " [k%d lpr%d r%d]", hn.ss.st[q] & 255, 1 << ((hn.ss.st[q] >> 8) & 255),
                  hn.ss.st[q + 1] & 0xffff);
  if (o < (int)sizeof(buf) - 64) o += snprintf(buf + o, sizeof(buf) - o, "\nrank-1 levels (rows,cols):");
  for (size_t q = 0; q < hn.glu.r1.size() && o < (int)sizeof(buf) - 64; q += 8)
    o += snprintf(buf + o, sizeof(buf) - o, " (%d,%d)", hn.glu.r1[q + 5], 4 * hn.glu.r1[q + 6]);
  snprintf(buf + o, sizeof(buf) - o, "\n");
  return std::string(buf);
}

// Host emulation of the stand-alone Jacobian kernel's tables (HostNet::JacCols incl. the repacked
// pairw / slotw / ent4 / kind pointers of jac_kernel_pipe) against a direct evaluation of
// chem_ode_jac (src/disk.f90:4746-4903, evolT = F) on one pseudo-random state.
static bool selfcheck_jac_tables(const HostNet& hn, std::string& err) {
  const HostNet::JacCols& jc = hn.jc;
  auto fail = [&](const std::string& m) { err = "K3 tables: " + m; return false; };
  const int R = hn.R, NEQ = hn.NEQ, NNZ = hn.NNZ;
  if (jc.pairw.size() != 2 * jc.pair.size() || jc.slotw.size() != 2 * jc.slot_id.size() ||
      (int)jc.grp_two_ptr.size() != jc.ngroups || (int)jc.grp_sat_ptr.size() != jc.ngroups || jc.ent4.size() % 4)
    return fail("sizes");
  std::vector<double> k(R), y(NEQ);
  uint64_t st = 88172645463325252ull;
  auto rnd = [&] { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0; };
  for (auto& v : k) v = 0.5 + rnd();
  for (auto& v : y) { v = 0.5 + rnd(); if (rnd() < 0.05) v = -v; }
  const double DS = 1.7;
  auto slot_of = [&](int row, int col) -> int {
    for (int q = hn.ia[col] - 1; q < hn.ia[col + 1] - 1; ++q) if (hn.ja[q] - 1 == row) return q;
    return -1;
  };
  std::vector<double> ref(NNZ, 0.0), scale(NNZ, 0.0), tab(NNZ, 0.0);
  for (int i = 0; i < R; ++i) {
    const uint32_t w = hn.fw[i];
    const int kind = (w >> 20) & 3, r1 = w & 1023, r2 = (w >> 10) & 1023;
    if (kind == FK_SKIP) continue;
    std::map<int, int> net;
    for (int q = 0; q < hn.n_reac[i]; ++q) net[hn.reac[3 * i + q] - 1] -= 1;
    for (int q = 0; q < hn.n_prod[i]; ++q) net[hn.prod[4 * i + q] - 1] += 1;
    for (int which = 0; which < 2; ++which) {
      double d;
      int col;
      if (kind == FK_ONE) { if (which) continue; d = k[i]; col = r1; }
      else if (kind == FK_TWO) {
        if (which && r2 == r1) continue;
        d = (r1 != r2) ? (which == 0 ? k[i] * y[r2] : k[i] * y[r1]) : 2.0 * k[i] * y[r2];
        if (y[r1] < 0.0 && y[r2] < 0.0) d = -d;
        col = which == 0 ? r1 : r2;
      } else {
        if (which) continue;
        const double tmp2 = DS * hn.sat_c[w >> 22];
        d = 0.0;
        if (tmp2 > 0.0) { const double tmp1 = 1.0 / tmp2, tmp = y[r1] * tmp1; d = (tmp <= 1e-4) ? k[i] * tmp1 : k[i] * tmp1 * std::exp(-tmp); }
        col = r1;
      }
      for (auto& kv : net) {
        if (kv.second == 0) continue;
        const int sl = slot_of(kv.first, col);
        if (sl < 0) return fail("a reaction term falls outside the CSC pattern");
        ref[sl] += kv.second * d; scale[sl] += std::fabs(kv.second * d);
      }
    }
  }
  std::vector<char> seen(NNZ, 0);
  std::vector<double> dbuf((size_t)jc.max_pairs + 1, 0.0);
  for (int g = 0; g < jc.ngroups; ++g) {
    const int pb = jc.grp_pair_ptr[g], pe = jc.grp_pair_ptr[g + 1], p2 = jc.grp_two_ptr[g], p3 = jc.grp_sat_ptr[g];
    if (!(pb <= p2 && p2 <= p3 && p3 <= pe) || pe - pb > jc.max_pairs) return fail("group pointers");
    for (int p = pb; p < pe; ++p) {
      const uint32_t x = jc.pairw[2 * p], r = jc.pairw[2 * p + 1];
      if (r != (jc.pair[p] & 0xffffu) || (int)r >= R) return fail("pair word / reaction");
      const int kind = (hn.fw[r] >> 20) & 3;
      if (kind != (p < p2 ? FK_ONE : p < p3 ? FK_TWO : FK_SAT)) return fail("pairs are not sorted by kind");
      double d;
      if (p < p2) d = k[r];
      else if (p < p3) {
        const double yo = y[x & 1023u], ys = y[(x >> 10) & 1023u];
        d = ((x >> 20) ? 2.0 : 1.0) * k[r] * yo;
        if (yo < 0.0 && ys < 0.0) d = -d;
      } else {
        const double tmp2 = DS * hn.sat_c[x >> 10];
        d = 0.0;
        if (tmp2 > 0.0) { const double tmp1 = 1.0 / tmp2, tmp = y[x & 1023u] * tmp1; d = (tmp <= 1e-4) ? k[r] * tmp1 : k[r] * tmp1 * std::exp(-tmp); }
      }
      dbuf[p - pb] = d;
    }
    for (int q = pe - pb; q < jc.max_pairs; ++q) dbuf[q] = std::nan("");   // stale rows must never be read
    dbuf[jc.max_pairs] = 0.0;
    for (int s2 = jc.grp_slot_ptr[g]; s2 < jc.grp_slot_ptr[g + 1]; ++s2) {
      const uint32_t slot = jc.slotw[2 * s2], off = jc.slotw[2 * s2 + 1] & 0xffffffu, n4 = jc.slotw[2 * s2 + 1] >> 24;
      if ((int)slot >= NNZ || n4 < 1 || (size_t)4 * (off + n4) > jc.ent4.size()) return fail("slot word");
      if (!jc.grp_accum[g] && seen[slot]) return fail("a slot is stored twice");
      if (jc.grp_accum[g] && !seen[slot] && std::find(jc.zero_slots.begin(), jc.zero_slots.end(), (int)slot) == jc.zero_slots.end())
        return fail("an accumulating chunk meets a slot nobody initialised");
      double acc = jc.grp_accum[g] ? tab[slot] : 0.0;
      for (size_t e = (size_t)4 * off; e < (size_t)4 * (off + n4); ++e) {
        const uint32_t v = jc.ent4[e], byte = v & 0x1ffffu;
        if (byte % 512u || byte / 512u > (uint32_t)jc.max_pairs) return fail("entry offset");
        const uint64_t bits = (uint64_t)(v & 0xfffe0000u) << 32;
        double cf; memcpy(&cf, &bits, 8);
        acc += cf * dbuf[byte / 512u];
      }
      tab[slot] = acc; seen[slot] = 1;
    }
  }
  for (int z : jc.zero_slots) if (z < 0 || z >= NNZ) return fail("zero slot index");
  for (int q = 0; q < NNZ; ++q) {
    const bool zero = std::find(jc.zero_slots.begin(), jc.zero_slots.end(), q) != jc.zero_slots.end();
    if (!seen[q] && !zero) return fail("a CSC slot is neither gathered nor zeroed");
    if (!(std::fabs(tab[q] - ref[q]) <= 1e-12 * scale[q])) return fail("gathered value differs from the direct evaluation at slot " + std::to_string(q));
  }
  return true;
}

bool selfcheck_schedules(const HostNet& hn, std::string& err) {
  const HostNet::LevelLU& g = hn.glu;
  const HostNet::SolveSched& ss = hn.ss;
  if (!selfcheck_jac_tables(hn, err)) return false;
  if (g.nlev == 0) { err = "no level-parallel schedule for this network"; return false; }
  const int nh = hn.nh, nt = hn.nt;
  auto fail = [&](const std::string& m) { err = m; return false; };
  // storage position of (i,j) in the permuted numbering, -1 if outside the pattern
  auto pos = [&](int i, int j) -> int {
    if (i >= nh && j >= nh) return hn.o_tl + (j - nh) * hn.ldt + (i - nh);
    if (i < nh && j < nh) {
      auto lo = hn.hh_col.begin() + hn.hh_ptr[i], hi = hn.hh_col.begin() + hn.hh_ptr[i + 1];
      auto it = std::lower_bound(lo, hi, (uint16_t)j);
      return (it == hi || *it != j) ? -1 : (int)(it - hn.hh_col.begin());
    }
    if (i < nh) {
      auto lo = hn.ub_col.begin() + hn.ub_ptr[i], hi = hn.ub_col.begin() + hn.ub_ptr[i + 1];
      auto it = std::lower_bound(lo, hi, (uint16_t)(j - nh));
      return (it == hi || *it != j - nh) ? -1 : hn.o_ub + (int)(it - hn.ub_col.begin());
    }
    auto lo = hn.lc_col.begin() + hn.lc_ptr[i - nh], hi = hn.lc_col.begin() + hn.lc_ptr[i - nh + 1];
    auto it = std::lower_bound(lo, hi, (uint16_t)j);
    return (it == hi || *it != j) ? -1 : hn.o_lc + (int)(it - hn.lc_col.begin());
  };
  // rows below / columns right of every head pivot
  std::vector<std::vector<int>> rows_of(nh), cols_of(nh);
  for (int i = 0; i < nh; ++i) {
    for (int q = hn.hh_ptr[i]; q < hn.hh_ptr[i] + hn.hh_nl[i]; ++q) rows_of[hn.hh_col[q]].push_back(i);
    for (int q = hn.hh_ptr[i] + hn.hh_nl[i] + 1; q < hn.hh_ptr[i + 1]; ++q) cols_of[i].push_back(hn.hh_col[q]);
    for (int q = hn.ub_ptr[i]; q < hn.ub_ptr[i + 1]; ++q) cols_of[i].push_back(nh + hn.ub_col[q]);
    if (hn.hh_col[hn.hh_ptr[i] + hn.hh_nl[i]] != i) return fail("diagonal entry misplaced in a head row");
  }
  for (int a = 0; a < nt; ++a)
    for (int q = hn.lc_ptr[a]; q < hn.lc_ptr[a + 1]; ++q) rows_of[hn.lc_col[q]].push_back(nh + a);
  // level of every pivot from the schedule
  std::vector<int> level_of(nh, -1);
  for (int L = 0; L < g.nlev; ++L)
    for (int q = g.lvl[4 * L]; q < g.lvl[4 * (L + 1)]; ++q) {
      const int k = (int)(g.piv[q] >> 16);
      if (k < 0 || k >= nh || level_of[k] >= 0) return fail("pivot listed twice or out of range");
      if ((int)(g.piv[q] & 0xffff) != pos(k, k)) return fail("wrong diagonal position of a pivot");
      level_of[k] = L;
    }
  for (int k = 0; k < nh; ++k) if (level_of[k] < 0) return fail("pivot missing from the level schedule");
  long npairs = 0;
  for (int L = 0; L < g.nlev; ++L) {
    // expected work of the level
    std::map<std::pair<int, int>, int> expect;            // (l position, u position) -> target
    std::set<std::pair<int, int>> expect_mul;             // (position of a(i,k), diagonal position)
    std::set<int> operands;
    for (int k = 0; k < nh; ++k) {
      if (level_of[k] != L) continue;
      for (int i : rows_of[k]) {
        const int pl = pos(i, k);
        expect_mul.insert({pl, pos(k, k)});
        operands.insert(pl);
        for (int j : cols_of[k]) {
          const int pu = pos(k, j), pt = pos(i, j);
          if (pl < 0 || pu < 0 || pt < 0) return fail("update outside the symbolic pattern");
          expect[{pl, pu}] = pt;
          operands.insert(pu);
        }
      }
    }
    // multipliers
    std::set<std::pair<int, int>> got_mul;
    for (int q = g.lvl[4 * L + 1]; q < g.lvl[4 * (L + 1) + 1]; ++q)
      if (!got_mul.insert({(int)(g.mul[q] & 0xffff), (int)(g.mul[q] >> 16)}).second) return fail("multiplier listed twice");
    if (got_mul != expect_mul) return fail("multiplier list of a level differs from the pattern");
    // updates
    std::map<std::pair<int, int>, int> got;
    std::set<int> targets;
    const int r1 = g.lvl[4 * L + 3];
    if (r1 >= 0) {
      const int* A = &g.r1[8 * r1];
      const int ua0 = A[0], nua4 = A[1], ub0 = A[2], tgt_off = A[4], nr = A[5], nj4 = A[6];
      const int m0 = g.lvl[4 * L + 1];
      for (int rc = 0; rc * 32 < nr; ++rc)
        for (int j4 = 0; j4 < nj4; ++j4)
          for (int l = 0; l < 32; ++l)
            for (int c = 0; c < 4; ++c) {
              const int t = g.r1tgt[((size_t)tgt_off + ((size_t)rc * nj4 + j4) * 32 + l) * 4 + c];
              if (t == g.zpos + 1) continue;
              const int m = rc * 32 + l;
              if (m >= nr) return fail("rank-1 level: live target in an idle lane");
              const int pl = (int)(g.mul[m0 + m] & 0xffff);
              const int pu = (j4 < nua4) ? ua0 + 4 * j4 + c : ub0 + 4 * (j4 - nua4) + c;
              if (!got.insert({{pl, pu}, t}).second) return fail("rank-1 level: update listed twice");
              if (!targets.insert(t).second) return fail("rank-1 level: target written twice");
            }
    }
    for (int gi = g.lvl[4 * L + 2]; gi < g.lvl[4 * (L + 1) + 2]; ++gi) {
      const int width = g.grp[4 * gi], nblk = g.grp[4 * gi + 1], ent_off = g.grp[4 * gi + 2], tgt_off = g.grp[4 * gi + 3];
      for (int b = 0; b < nblk; ++b)
        for (int l = 0; l < 32; ++l) {
          const int t = g.tgt[(size_t)tgt_off + b * 32 + l];
          bool any = false;
          for (int j = 0; j < width; ++j) {
            const uint32_t e = g.ent[(size_t)ent_off + ((size_t)b * width + j) * 32 + l];
            const int pl = (int)(e & 0xffff), pu = (int)(e >> 16);
            if (pl == g.zpos && pu == g.zpos) continue;
            if (t == 0xFFFF) return fail("live entry in an idle lane");
            if (!got.insert({{pl, pu}, t}).second) return fail("update listed twice");
            any = true;
          }
          if (any && !targets.insert(t).second) return fail("target written by two lanes of one level");
        }
    }
    if (got != expect) return fail("updates of a level differ from the symbolic factorisation (level " + std::to_string(L) + ")");
    for (int t : targets) if (operands.count(t)) return fail("a target of a level is also an operand of that level");
    npairs += (long)got.size();
  }
  if (npairs != g.npairs) return fail("pair count mismatch");
  // staged solves: every L / U entry of the head block exactly once (stage entry or S diagonal block)
  std::vector<int> seenL(hn.n_hh, 0), seenU(hn.n_hh, 0);
  const uint32_t* ent = ss.blob.data();
  const uint16_t* rp = (const uint16_t*)(ent + ss.nent);
  const uint16_t* rows = rp + ((ss.nrp + 1) & ~1);
  const int nstage = (int)ss.st.size() / 4;
  if (nstage != ss.nf + ss.nb) return fail("stage count mismatch");
  std::vector<int> done_f(nh, 0), done_b(nh, 0);
  for (int s = 0; s < nstage; ++s) {
    const bool upper = s >= ss.nf;
    const int kind = ss.st[4 * s] & 255, nrows = ss.st[4 * s + 1] & 0xffff, row_off = ss.st[4 * s + 2], rp_off = ss.st[4 * s + 3];
    if ((kind == 1) != (upper && kind != 2) || (kind == 0 && upper)) return fail("stage kind does not match its sweep");
    for (int r = 0; r < nrows; ++r) {
      const int i = rows[row_off + r];
      for (int q = rp[rp_off + r]; q < rp[rp_off + r + 1]; ++q) {
        const int p = (int)(ent[q] & 0xffff), c = (int)(ent[q] >> 16);
        if (p < hn.hh_ptr[i] || p >= hn.hh_ptr[i + 1] || hn.hh_col[p] != c) return fail("stage entry does not belong to its row");
        if (upper ? (c <= i) : (c >= i)) return fail("stage entry on the wrong side of the diagonal");
        // the operand must already be solved: an earlier stage of the same sweep
        if (!(upper ? done_b[c] : done_f[c]) && !(!upper && hn.hh_nl[c] == 0)) return fail("stage reads a row that is not solved yet");
        (upper ? seenU : seenL)[p] += 1;
      }
    }
    for (int r = 0; r < nrows; ++r) (upper ? done_b : done_f)[rows[row_off + r]] = 1;
  }
  for (uint32_t e : ss.ext) {
    const int p = (int)(e & 0xffff);
    if (p < 0 || p >= hn.n_hh) return fail("extraction entry outside the head block");
    seenL[p] += 1; seenU[p] += 1;       // diagonal-block entries are handled by the block solve
  }
  for (int i = 0; i < nh; ++i)
    for (int q = hn.hh_ptr[i]; q < hn.hh_ptr[i + 1]; ++q) {
      const int c = hn.hh_col[q];
      if (c < i && seenL[q] != 1) return fail("an L entry of the head block is not covered exactly once by the forward stages");
      if (c > i && seenU[q] != 1) return fail("a U entry of the head block is not covered exactly once by the backward stages");
    }
  return true;
}

}  // namespace racg
