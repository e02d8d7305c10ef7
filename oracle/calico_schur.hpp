// ORACLE — TEST INFRASTRUCTURE ONLY (see jet.hpp header).
//
// Linear solvers for the LM step (J^T J + D^2) y = J^T r.
//   0  dense normal equations + Cholesky (the trusted, structure-free path)
//   1  Schur complement with the spline control points as e-blocks in time order: block-banded
//      Cholesky (the banded structure the reference notes but does not exploit, bspline.hpp:287-289)
//   2  Ceres-style DENSE_SCHUR with automatic ordering (batch_optimizer.cpp:12): e-blocks are a
//      maximal independent set of parameter blocks, the reduced system is dense.
//      Used as the timed "restated Ceres" CPU baseline.
// All three return the same step up to rounding.
#pragma once
#include "calico_lm.hpp"

namespace orc {

// Banded Cholesky of an n x n SPD matrix with half-bandwidth hb stored as rows of (hb+1):
// Ab[i*(hb+1) + (hb - (i-j))] = A(i,j) for i-hb <= j <= i. In place → L in the same layout.
inline bool BandedCholesky(double* Ab, int n, int hb) {
  const int w = hb + 1;
  for (int j = 0; j < n; ++j) {
    double d = Ab[size_t(j) * w + hb];
    for (int t = std::max(0, j - hb); t < j; ++t) { const double l = Ab[size_t(j) * w + hb - (j - t)]; d -= l * l; }
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    d = std::sqrt(d);
    Ab[size_t(j) * w + hb] = d;
    const int imax = std::min(n - 1, j + hb);
    for (int i = j + 1; i <= imax; ++i) {
      double s = Ab[size_t(i) * w + hb - (i - j)];
      for (int t = std::max(0, i - hb); t < j; ++t) s -= Ab[size_t(i) * w + hb - (i - t)] * Ab[size_t(j) * w + hb - (j - t)];
      Ab[size_t(i) * w + hb - (i - j)] = s / d;
    }
  }
  return true;
}
// Solve L X = B in place for nrhs right-hand sides; B is n x nrhs row-major.
inline void BandedForward(const double* Lb, int n, int hb, double* B, int nrhs) {
  const int w = hb + 1;
  for (int i = 0; i < n; ++i) {
    double* bi = B + size_t(i) * nrhs;
    for (int t = std::max(0, i - hb); t < i; ++t) { const double l = Lb[size_t(i) * w + hb - (i - t)]; const double* bt = B + size_t(t) * nrhs; for (int c = 0; c < nrhs; ++c) bi[c] -= l * bt[c]; }
    const double inv = 1.0 / Lb[size_t(i) * w + hb];
    for (int c = 0; c < nrhs; ++c) bi[c] *= inv;
  }
}
inline void BandedBackward(const double* Lb, int n, int hb, double* b) {
  const int w = hb + 1;
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    const int tmax = std::min(n - 1, i + hb);
    for (int t = i + 1; t <= tmax; ++t) s -= Lb[size_t(t) * w + hb - (t - i)] * b[t];
    b[i] = s / Lb[size_t(i) * w + hb];
  }
}

inline bool SolveBandedSchur(Minimizer& M, const std::vector<double>& D, std::vector<double>& y) {
  const int na = M.n_cp_tan, nc = M.n_tan - na, n = M.n_tan;
  const int hb = 6 * M.p.k - 1;
  const int w = hb + 1;
  std::vector<double> Ab(size_t(na) * w, 0.0), B(size_t(na) * (nc + 1), 0.0), C(size_t(nc) * nc, 0.0), gc(nc, 0.0);
  // B carries g_a as its last column.
  for (size_t ai = 0; ai < M.tiles.size(); ++ai) {
    const int m = M.p.rblocks[M.active_rblocks[ai]].m;
    const auto& tl = M.tiles[ai];
    const double* r = &M.residuals[M.row0[ai]];
    for (const auto& a : tl) {
      const double* Ja = M.jvals.data() + a.val;
      for (int i = 0; i < a.t; ++i) {
        double gi = 0; for (int q = 0; q < m; ++q) gi += Ja[q * a.t + i] * r[q];
        if (a.off < na) B[size_t(a.off + i) * (nc + 1) + nc] += gi; else gc[a.off - na + i] += gi;
      }
      for (const auto& b : tl) {
        const double* Jb = M.jvals.data() + b.val;
        for (int i = 0; i < a.t; ++i) for (int j = 0; j < b.t; ++j) {
          const int gi = a.off + i, gj = b.off + j;
          double s = 0; for (int q = 0; q < m; ++q) s += Ja[q * a.t + i] * Jb[q * b.t + j];
          if (gi < na && gj < na) { if (gj <= gi) Ab[size_t(gi) * w + hb - (gi - gj)] += s; }
          else if (gi < na && gj >= na) B[size_t(gi) * (nc + 1) + (gj - na)] += s;
          else if (gi >= na && gj >= na) C[size_t(gi - na) * nc + (gj - na)] += s;
        }
      }
    }
  }
  for (int i = 0; i < na; ++i) Ab[size_t(i) * w + hb] += D[i] * D[i];
  for (int i = 0; i < nc; ++i) C[size_t(i) * nc + i] += D[na + i] * D[na + i];
  if (!BandedCholesky(Ab.data(), na, hb)) return false;
  BandedForward(Ab.data(), na, hb, B.data(), nc + 1);  // W | z
  // S = C - W^T W ; rhs = g_c - W^T z.
  for (int t = 0; t < na; ++t) {
    const double* wt = &B[size_t(t) * (nc + 1)];
    for (int i = 0; i < nc; ++i) { const double wi = wt[i]; if (wi == 0.0) continue; for (int j = 0; j <= i; ++j) C[size_t(i) * nc + j] -= wi * wt[j]; }
    for (int i = 0; i < nc; ++i) gc[i] -= wt[i] * wt[nc];
  }
  if (nc > 0) { if (!DenseCholesky(C.data(), nc)) return false; CholeskySolve(C.data(), nc, gc.data()); }
  y.assign(n, 0.0);
  for (int i = 0; i < nc; ++i) y[na + i] = gc[i];
  for (int t = 0; t < na; ++t) { const double* wt = &B[size_t(t) * (nc + 1)]; double s = wt[nc]; for (int i = 0; i < nc; ++i) s -= wt[i] * gc[i]; y[t] = s; }
  BandedBackward(Ab.data(), na, hb, y.data());
  return true;
}

// Ceres-style DENSE_SCHUR with automatic ordering. Ceres's ComputeStableSchurOrdering / independent-set
// ordering (parameter_block_ordering.cc) greedily picks parameter blocks none of which share a residual
// block; for this problem family that selects every k-th control point (and nothing else once the
// sensors' blocks are adjacent to every control point), leaving a dense reduced system over the other
// control points + calibration blocks. Here the independent set is built greedily in block order.
inline bool SolveCeresDenseSchur(Minimizer& M, const std::vector<double>& D, std::vector<double>& y) {
  Problem& p = M.p;
  const int n = M.n_tan;
  // Greedy maximal independent set over active parameter blocks, lowest degree first is what Ceres
  // does; control points all have equal degree order, so in-order greedy reproduces "every k-th".
  std::vector<std::vector<int>> adj_rb(p.blocks.size());
  for (size_t ai = 0; ai < M.tiles.size(); ++ai) for (const auto& t : M.tiles[ai]) adj_rb[t.block].push_back(int(ai));
  std::vector<char> state(p.blocks.size(), 0);  // 0 white, 1 e-block, 2 excluded
  std::vector<int> order = M.active_blocks;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return adj_rb[a].size() < adj_rb[b].size(); });
  for (int id : order) {
    if (state[id]) continue;
    state[id] = 1;
    for (int ai : adj_rb[id]) for (const auto& t : M.tiles[ai]) if (!state[t.block]) state[t.block] = 2;
  }
  // New ordering: e-blocks first.
  std::vector<int> newoff(p.blocks.size(), -1);
  int ne = 0;
  std::vector<int> eblocks;
  for (int id : M.active_blocks) if (state[id] == 1) { newoff[id] = ne; ne += p.blocks[id].tsize; eblocks.push_back(id); }
  int nf = 0;
  for (int id : M.active_blocks) if (state[id] != 1) { newoff[id] = ne + nf; nf += p.blocks[id].tsize; }
  std::vector<double> S(size_t(nf) * nf, 0.0), rhs(nf, 0.0);
  // f-f part of J^T J and g_f from all rows.
  for (size_t ai = 0; ai < M.tiles.size(); ++ai) {
    const int m = p.rblocks[M.active_rblocks[ai]].m;
    const double* r = &M.residuals[M.row0[ai]];
    for (const auto& a : M.tiles[ai]) {
      if (state[a.block] == 1) continue;
      const double* Ja = M.jvals.data() + a.val; const int oa = newoff[a.block] - ne;
      for (int i = 0; i < a.t; ++i) { double g = 0; for (int q = 0; q < m; ++q) g += Ja[q * a.t + i] * r[q]; rhs[oa + i] += g; }
      for (const auto& b : M.tiles[ai]) {
        if (state[b.block] == 1) continue;
        const int ob = newoff[b.block] - ne; if (ob > oa) continue;
        const double* Jb = M.jvals.data() + b.val;
        for (int i = 0; i < a.t; ++i) for (int j = 0; j < b.t; ++j) { double s = 0; for (int q = 0; q < m; ++q) s += Ja[q * a.t + i] * Jb[q * b.t + j]; S[size_t(oa + i) * nf + ob + j] += s; }
      }
    }
  }
  for (int id : M.active_blocks) if (state[id] != 1) for (int i = 0; i < p.blocks[id].tsize; ++i) { const double d = D[p.blocks[id].off + i]; S[size_t(newoff[id] - ne + i) * nf + newoff[id] - ne + i] += d * d; }
  // Eliminate each e-block (SchurEliminator::Eliminate chunk by chunk).
  struct EData { std::vector<double> Linv_E_T; };
  std::vector<std::vector<double>> e_inv(eblocks.size()), e_g(eblocks.size());
  std::vector<std::vector<int>> e_fblocks(eblocks.size());
  std::vector<std::vector<double>> e_EF(eblocks.size());  // E^T F per f-block, concatenated
  for (size_t e = 0; e < eblocks.size(); ++e) {
    const int id = eblocks[e]; const int te = p.blocks[id].tsize;
    std::vector<double> EtE(size_t(te) * te, 0.0), g(te, 0.0);
    std::vector<int> fb; std::vector<int> fboff;  // f-blocks in this chunk
    int ftot = 0;
    for (int ai : adj_rb[id]) for (const auto& t : M.tiles[ai]) if (state[t.block] != 1 && std::find(fb.begin(), fb.end(), t.block) == fb.end()) { fb.push_back(t.block); fboff.push_back(ftot); ftot += t.t; }
    std::vector<double> EtF(size_t(te) * ftot, 0.0);
    for (int ai : adj_rb[id]) {
      const int m = p.rblocks[M.active_rblocks[ai]].m;
      const double* r = &M.residuals[M.row0[ai]];
      const double* E = nullptr;
      for (const auto& t : M.tiles[ai]) if (t.block == id) E = M.jvals.data() + t.val;
      for (int i = 0; i < te; ++i) { for (int q = 0; q < m; ++q) g[i] += E[q * te + i] * r[q]; for (int j = 0; j < te; ++j) { double s = 0; for (int q = 0; q < m; ++q) s += E[q * te + i] * E[q * te + j]; EtE[size_t(i) * te + j] += s; } }
      for (const auto& t : M.tiles[ai]) {
        if (state[t.block] == 1) continue;
        const int fo = fboff[std::find(fb.begin(), fb.end(), t.block) - fb.begin()];
        const double* F = M.jvals.data() + t.val;
        for (int i = 0; i < te; ++i) for (int j = 0; j < t.t; ++j) { double s = 0; for (int q = 0; q < m; ++q) s += E[q * te + i] * F[q * t.t + j]; EtF[size_t(i) * ftot + fo + j] += s; }
      }
    }
    for (int i = 0; i < te; ++i) { const double d = D[p.blocks[id].off + i]; EtE[size_t(i) * te + i] += d * d; }
    if (!DenseCholesky(EtE.data(), te)) return false;
    // W = L^-1 EtF, z = L^-1 g ; S -= W^T W ; rhs -= W^T z.
    for (int c = 0; c < ftot; ++c) { for (int i = 0; i < te; ++i) { double s = EtF[size_t(i) * ftot + c]; for (int t = 0; t < i; ++t) s -= EtE[size_t(i) * te + t] * EtF[size_t(t) * ftot + c]; EtF[size_t(i) * ftot + c] = s / EtE[size_t(i) * te + i]; } }
    for (int i = 0; i < te; ++i) { double s = g[i]; for (int t = 0; t < i; ++t) s -= EtE[size_t(i) * te + t] * g[t]; g[i] = s / EtE[size_t(i) * te + i]; }
    // Map chunk-local f columns to reduced offsets.
    std::vector<int> col(ftot);
    for (size_t b = 0; b < fb.size(); ++b) for (int j = 0; j < p.blocks[fb[b]].tsize; ++j) col[fboff[b] + j] = newoff[fb[b]] - ne + j;
    for (int a = 0; a < ftot; ++a) {
      double dr = 0; for (int i = 0; i < te; ++i) dr += EtF[size_t(i) * ftot + a] * g[i];
      rhs[col[a]] -= dr;
      for (int b = 0; b < ftot; ++b) { if (col[b] > col[a]) continue; double s = 0; for (int i = 0; i < te; ++i) s += EtF[size_t(i) * ftot + a] * EtF[size_t(i) * ftot + b]; S[size_t(col[a]) * nf + col[b]] -= s; }
    }
    e_inv[e] = std::move(EtE); e_g[e] = std::move(g); e_fblocks[e] = std::move(fb); e_EF[e] = std::move(EtF);
  }
  if (nf > 0) { if (!DenseCholesky(S.data(), nf)) return false; CholeskySolve(S.data(), nf, rhs.data()); }
  y.assign(n, 0.0);
  for (int id : M.active_blocks) if (state[id] != 1) for (int i = 0; i < p.blocks[id].tsize; ++i) y[p.blocks[id].off + i] = rhs[newoff[id] - ne + i];
  // Back-substitute: y_e = L^-T (z - W y_f).
  for (size_t e = 0; e < eblocks.size(); ++e) {
    const int id = eblocks[e]; const int te = p.blocks[id].tsize;
    const auto& fb = e_fblocks[e]; int ftot = 0; for (int b : fb) ftot += p.blocks[b].tsize;
    std::vector<double> v = e_g[e];
    int fo = 0;
    for (int b : fb) { for (int j = 0; j < p.blocks[b].tsize; ++j) { const double yf = y[p.blocks[b].off + j]; for (int i = 0; i < te; ++i) v[i] -= e_EF[e][size_t(i) * ftot + fo + j] * yf; } fo += p.blocks[b].tsize; }
    const double* L = e_inv[e].data();
    for (int i = te - 1; i >= 0; --i) { double s = v[i]; for (int t = i + 1; t < te; ++t) s -= L[size_t(t) * te + i] * v[t]; v[i] = s / L[size_t(i) * te + i]; }
    for (int i = 0; i < te; ++i) y[p.blocks[id].off + i] = v[i];
  }
  return true;
}

inline bool Minimizer::LinearSolve(const std::vector<double>& D, std::vector<double>& y) {
  switch (opt.linear_solver) {
    case 1: return SolveBandedSchur(*this, D, y);
    case 2: return SolveCeresDenseSchur(*this, D, y);
    default: return SolveDenseNormal(D, y);
  }
}

}  // namespace orc
