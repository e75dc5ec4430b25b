// raco_sparse.cpp -- ORACLE (test infrastructure only, see raco.h): fixed-pattern
// sparse LU WITHOUT pivoting, the role YSMP plays for DLSODES (src/opkda1.f:
// ODRV/md 1994-2714 minimum-degree ordering of M+M^T, CDRV/nroc/nsfc 2715-3542
// symbolic factorisation, nnfc 3543-3697 row-wise numeric LDU with no pivoting and
// a zero-pivot error flag, nntc 3750-3801 triangular solves).
// Restated at the level of the published algorithm: a minimum-degree ordering on
// the symmetrised pattern (ties -> lowest index; YSMP's md uses mass elimination
// and different tie-breaks, so the permutation, and the fill by a few entries, can
// differ -- the factorisation is the same mathematical object, P = L U exactly for
// the same P), then row-merge symbolic factorisation and a row-by-row numeric
// factorisation through a dense work row.
#include "raco_internal.hpp"
#include <algorithm>
#include <set>

namespace raco {

void SparseLU::analyse(int n_, const std::vector<int>& ia1, const std::vector<int>& ja1) {
  n = n_;
  // user's CSC (1-based): column c holds rows ja1[ia1[c]-1 .. ia1[c+1]-2]
  // Build row lists of A (with slots), adding missing diagonals (DPREP, src/opkda1.f:1372-1394)
  std::vector<std::vector<std::pair<int, int>>> rows(n);  // (col, slot)
  std::vector<char> has_diag(n, 0);
  for (int c = 0; c < n; ++c)
    for (int k = ia1[c] - 1; k < ia1[c + 1] - 1; ++k) {
      int r = ja1[k] - 1;
      rows[r].push_back({c, k});
      if (r == c) has_diag[c] = 1;
    }
  for (int i = 0; i < n; ++i) if (!has_diag[i]) rows[i].push_back({i, -1});
  nnz_a = 0;
  for (auto& r : rows) nnz_a += (int)r.size();
  // symmetrised adjacency for the ordering
  std::vector<std::set<int>> adj(n);
  for (int r = 0; r < n; ++r)
    for (auto& e : rows[r]) if (e.first != r) { adj[r].insert(e.first); adj[e.first].insert(r); }
  // minimum degree with explicit elimination graph
  perm.assign(n, 0); iperm.assign(n, 0);
  std::vector<char> done(n, 0);
  for (int step = 0; step < n; ++step) {
    int best = -1; size_t bd = (size_t)-1;
    for (int v = 0; v < n; ++v) if (!done[v] && adj[v].size() < bd) { bd = adj[v].size(); best = v; }
    done[best] = 1;
    perm[step] = best; iperm[best] = step;
    std::vector<int> nb(adj[best].begin(), adj[best].end());
    for (int a : nb) adj[a].erase(best);
    for (size_t x = 0; x < nb.size(); ++x)
      for (size_t y = x + 1; y < nb.size(); ++y) { adj[nb[x]].insert(nb[y]); adj[nb[y]].insert(nb[x]); }
    adj[best].clear();
  }
  // permuted CSR of A' = P A P^T
  a_ptr.assign(n + 1, 0); a_col.clear(); a_src.clear();
  for (int i = 0; i < n; ++i) {
    std::vector<std::pair<int, int>> r;
    for (auto& e : rows[perm[i]]) r.push_back({iperm[e.first], e.second});
    std::sort(r.begin(), r.end());
    for (auto& e : r) { a_col.push_back(e.first); a_src.push_back(e.second); }
    a_ptr[i + 1] = (int)a_col.size();
  }
  // symbolic: struct of row i of L (cols < i) and U (cols > i)
  l_ptr.assign(n + 1, 0); u_ptr.assign(n + 1, 0); l_col.clear(); u_col.clear();
  std::vector<char> mark(n);
  for (int i = 0; i < n; ++i) {
    std::fill(mark.begin(), mark.end(), 0);
    for (int k = a_ptr[i]; k < a_ptr[i + 1]; ++k) mark[a_col[k]] = 1;
    for (int k = 0; k < i; ++k) {
      if (!mark[k]) continue;
      for (int q = u_ptr[k]; q < u_ptr[k + 1]; ++q) mark[u_col[q]] = 1;
    }
    for (int k = 0; k < i; ++k) if (mark[k]) l_col.push_back(k);
    l_ptr[i + 1] = (int)l_col.size();
    for (int k = i + 1; k < n; ++k) if (mark[k]) u_col.push_back(k);
    u_ptr[i + 1] = (int)u_col.size();
  }
  nzl = (int)l_col.size(); nzu = (int)u_col.size();
  l_val.assign(nzl, 0.0); u_val.assign(nzu, 0.0); dinv.assign(n, 0.0);
}

int SparseLU::factor(const double* pval, double added_diag_value) {
  std::vector<double> w(n, 0.0);
  for (int i = 0; i < n; ++i) {
    for (int q = l_ptr[i]; q < l_ptr[i + 1]; ++q) w[l_col[q]] = 0.0;
    for (int q = u_ptr[i]; q < u_ptr[i + 1]; ++q) w[u_col[q]] = 0.0;
    w[i] = 0.0;
    for (int k = a_ptr[i]; k < a_ptr[i + 1]; ++k)
      w[a_col[k]] = (a_src[k] >= 0) ? pval[a_src[k]] : added_diag_value;
    for (int q = l_ptr[i]; q < l_ptr[i + 1]; ++q) {
      int k = l_col[q];
      double l = w[k] * dinv[k];
      l_val[q] = l;
      if (l != 0.0)
        for (int r = u_ptr[k]; r < u_ptr[k + 1]; ++r) w[u_col[r]] -= l * u_val[r];
    }
    double d = w[i];
    if (d == 0.0 || std::isnan(d)) return i + 1;  // zero pivot (nnfc flag 8n+k)
    dinv[i] = 1.0 / d;
    for (int q = u_ptr[i]; q < u_ptr[i + 1]; ++q) u_val[q] = w[u_col[q]];
  }
  return 0;
}

void SparseLU::solve(double* x) const {
  std::vector<double> z(n);
  for (int i = 0; i < n; ++i) z[i] = x[perm[i]];
  for (int i = 0; i < n; ++i) {
    double s = z[i];
    for (int q = l_ptr[i]; q < l_ptr[i + 1]; ++q) s -= l_val[q] * z[l_col[q]];
    z[i] = s;
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = z[i];
    for (int q = u_ptr[i]; q < u_ptr[i + 1]; ++q) s -= u_val[q] * z[u_col[q]];
    z[i] = s * dinv[i];
  }
  for (int i = 0; i < n; ++i) x[perm[i]] = z[i];
}

}  // namespace raco
