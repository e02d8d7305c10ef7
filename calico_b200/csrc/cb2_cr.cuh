// calico_b200 — K5 level 1 by BLOCK CYCLIC REDUCTION: elimination of the spline control points of one chunk.
//
// A control point couples only to the k-1 = 5 control points on either side (camera_cost_functor.cpp:52-60), so in blocks of
// 5 control points (30 unknowns) the control-point Hessian A of a chunk is BLOCK TRIDIAGONAL, with a dense border
// [left separator 30 | right separator 30 | calibration N_c | rhs 1] (separators only with several chunks / GPUs).
// Odd-even (cyclic) reduction eliminates every other block of the still-active blocks per level: all eliminations of a level are
// independent, so the sequential depth is ceil(log2(n_blocks)) + 1 small dense steps (C4: 501 blocks -> 10 levels) instead of
// the 2 505 dependent column-block steps of a left-to-right banded Cholesky. It is the same Cholesky elimination in a
// nested-dissection ORDER (a symmetric permutation of an SPD matrix), i.e. the same LM step as Ceres's DENSE_SCHUR
// (batch_optimizer.cpp:12; Ceres external) up to rounding.
//
// One kernel per level, one CTA per ACTIVE block i (stride = 2^level, active blocks are the multiples of stride):
//   1. load the block's state  D_i (30x30), border_i (30 x nbw)  and subtract the pending Schur updates of the two blocks
//      eliminated next to it on the previous level (scratch U, below); level 0 gathers from the assembled normal equations.
//   2. survivor (even position): write the state back. Eliminated (odd position, or the last remaining block):
//        D_i = L L^T (6x6-blocked Cholesky in shared memory),
//        W = L^-1 [E_i | F_i | border_i]   (E_i, F_i: couplings to the active neighbours a = i - stride, b = i + stride;
//                                            one thread per column, L broadcast from shared memory),
//        U_i = [W_E W_F]^T W               (60 x (60 + nbw), FP64 tensor pipe: DMMA m8n8k4) -> scratch for the next level.
//      L_i, W_E|W_F are kept for the back-substitution; the border part of W lands in the chunk's W[n][nbw] array, which the
//      Gram / separator / calibration kernels of cb2_schur.cuh consume unchanged.
// Back-substitution, one kernel per level in reverse: x_i = L^-T (v_i - W_E x_a - W_F x_b), v = z - W_border y from
// border_matvec_kernel. One warp per block.
#pragma once
#include "cb2_normal.cuh"
#include "cb2_schur.cuh"

namespace cb2 {

constexpr int kCrB = 30;            // block size: (k-1) control points x 6
constexpr int kCrThreads = 256;
constexpr int kCrLs = 31;           // row stride of the 30x30 diagonal block in shared memory (odd: column walks of the Cholesky are conflict-free)
constexpr int kCrLc = 32;           // row stride of the finished factor's copy read by the triangular solve (16-byte broadcast loads)

CB2_HD int cr_row_stride(int nbw) { const int M = 2 * kCrB + nbw; return ((M + 7) / 8 * 8 - 4 + 15) / 16 * 16 + 4; }   // >= M rounded to 8, == 4 mod 16
CB2_HD size_t cr_smem_bytes(int nbw) { return (size_t(32) * cr_row_stride(nbw) + size_t(kCrB) * (kCrLs + kCrLc) + 34 + (nbw + 1) / 2) * sizeof(double); }
CB2_HD size_t cr_u_size(int nbw) { return size_t(2 * kCrB) * (2 * kCrB + nbw); }
// Staging area of the TMA load path of the non-first levels: [UL rows 30..59 | UR rows 0..29 | border rows of the block | diagonal block].
CB2_HD size_t cr_stage_doubles(int nbw) { return size_t(2 * kCrB) * (2 * kCrB + nbw) + size_t(kCrB) * nbw + size_t(kCrB) * kCrB; }
CB2_HD size_t cr_smem_bytes_tma(int nbw) { return (cr_smem_bytes(nbw) + 15) / 16 * 16 + cr_stage_doubles(nbw) * sizeof(double); }
CB2_HD int cr_levels(int nblk) { int l = 0; while (((nblk + (1 << l) - 1) >> l) > 1) ++l; return l + 1; }

#if defined(CB2_CR_CLOCKS) && !defined(CB2_EMUL)
#define CB2_CLK(k) do { if (threadIdx.x == 0) clk_[k] = clock64(); } while (0)
#else
#define CB2_CLK(k) do { } while (0)
#endif
constexpr int kCrBatch = 8;
// for e = t, t + T, ...: st(e, ld(e)), with kCrBatch loads in flight per thread.
template <class LoadF, class StoreF>
CB2_D void cr_batched(int total, int t, LoadF ld, StoreF st) {
  for (int e0 = t; e0 < total; e0 += kCrBatch * kCrThreads) {
    double v[kCrBatch];
#pragma unroll
    for (int u = 0; u < kCrBatch; ++u) v[u] = ld(min(e0 + u * kCrThreads, total - 1));   // clamped, never skipped: no branch between the loads
#pragma unroll
    for (int u = 0; u < kCrBatch; ++u) { const int e = e0 + u * kCrThreads; if (e < total) st(e, v[u]); }
  }
}

template <bool kFirst>
__global__ void __launch_bounds__(kCrThreads, (kFirst ? 2 : 1)) cr_level_kernel(const BandSys* __restrict__ systems, int level, long n_a, int N_c,
                                                             const double* __restrict__ Aband, const double* __restrict__ Bmat,
                                                             const double* __restrict__ Cmat, const double* __restrict__ grad,
                                                             const double* __restrict__ dtil2, double* __restrict__ scal, int use_tma,
                                                             int band_w = kCpCols, int b_stride = 0, int rhs_stride = 1, long src_row0 = -1) {
  // Where the first level reads the system from: the assembled normal equations (band of width band_w = 36 in Aband, border rows in Bmat
  // with stride b_stride = N_c, right-hand side grad, LM damping dtil2, rows src_row0 + i = the chunk's consecutive unknowns) or, for the
  // SEPARATOR level between the chunks, the reduced separator system the ranks have just summed (band_w = 60, border + rhs rows of
  // b_stride = N_c + 1 doubles, damping already added, src_row0 = 0).
  if (b_stride == 0) b_stride = N_c;
  const BandSys sy = systems[blockIdx.y];
  const int stride = 1 << level;
  const int j = blockIdx.x, i = j * stride;
  if (i >= sy.nblk || level >= cr_levels(sy.nblk)) return;   // chunks of unequal length: the shorter ones finish a level early
  const int nact = (sy.nblk + stride - 1) >> level;
  const bool elim = (j & 1) || nact == 1;
  const bool has_a = (j & 1) != 0;                              // active left neighbour i - stride (odd positions always have one)
  const bool has_b = (j & 1) && i + stride < sy.nblk;           // active right neighbour i + stride
  const int nbw = sy.nbw, n = sy.n, t = threadIdx.x;
  const int M = 2 * kCrB + nbw;                                 // columns of X = [E | F | border]
  const int XS = cr_row_stride(nbw);
  // Column split: gridDim.z CTAs work on one block. Everyone factors the (small) diagonal block and solves the 60 coupling columns [E | F]
  // — they are the left operand of the Schur update — but solves, multiplies and stores only its own slice of 8-column blocks of
  // X = [E | F | border]: the update U = [W_E W_F]^T W (80 % of the arithmetic, bound by the FP64 tensor pipe of ONE SM otherwise) and the
  // forward substitution spread over gridDim.z SMs. Slice 0 also owns the per-block outputs (factor, [W_E | W_F]^T, failure flag).
  const int nqb_all = (M + 7) / 8;
  const int zid = blockIdx.z, nz = gridDim.z;
  const int qb_lo = zid * nqb_all / nz, qb_hi = (zid + 1) * nqb_all / nz;
  const int sl_lo = max(8 * qb_lo, 2 * kCrB), sl_hi = min(M, 8 * qb_hi);      // own columns beyond the coupling columns (solved by everyone)
  const int n_extra = max(0, sl_hi - sl_lo);
  double* X = dyn_smem<double>();                               // [32][XS]; rows 30, 31 stay zero (k padding of the DMMA product)
  double* Dm = X + 32 * XS;                                     // [30][31]
  double* Lc = Dm + kCrB * kCrLs + (kCrB * kCrLs & 1);          // [30][32] copy of the factor, 16-byte aligned rows
  double* dinv = Lc + kCrB * kCrLc;                             // [30] reciprocal diagonal of L
  __shared__ int s_fail;
  const int r0 = i * kCrB;                                      // first chunk row of this block
  const size_t usz = cr_u_size(nbw);
  const int UW = M;                                             // row length of a scratch entry
  // Scratch of the blocks eliminated on the previous level to the left / right of block i.
  const double* UL = nullptr;
  const double* UR = nullptr;
  if (!kFirst) {
    const int h = stride >> 1;
    const double* Uprev = sy.crU + size_t((level - 1) & 1) * sy.cr_uslots * usz;
    if (i - h >= 0) UL = Uprev + size_t((i - h) / stride) * usz;
    if (i + h < sy.nblk) UR = Uprev + size_t((i + h) / stride) * usz;
  }
  long long clk_[8];
  (void)clk_;
  CB2_CLK(0);
  if (t == 0) s_fail = 0;
  // Global loads are issued in batches of kCrBatch independent loads per thread before their first use: the level kernels are
  // latency-bound (one CTA per block, ~17 k doubles from L2), so memory-level parallelism is what sets their duration.
  const double* __restrict__ crD = sy.crD + size_t(i) * (kCrB * kCrB);
  const double* __restrict__ crBd = sy.crBd + size_t(i) * kCrB * nbw;
  const long g0 = kFirst ? (src_row0 >= 0 ? src_row0 : long(sy.row_gidx[0])) : 0;   // level-1 chunk rows are consecutive unknowns: row -> g0 + row
  int* colg = reinterpret_cast<int*>(dinv + 32);                 // [nbw] global unknown of every border column (level 0 only)
  if (kFirst) {
    for (int c = t; c < nbw - 1; c += kCrThreads) colg[c] = sy.col_gidx[c];
    __syncthreads();
  }
  // Every value below is ONE unconditional load from a clamped / substituted address followed by selects, so that the kCrBatch
  // loads of a batch really are in flight together (no branches between them). Absent neighbours read the all-zero slot crZero.
  const double* __restrict__ ULz = UL ? UL : sy.crZero;
  const double* __restrict__ URz = UR ? UR : sy.crZero;
  auto band_ptr = [&](long gi, long gj, bool& ok) -> const double* {   // &A(gi, gj) of the assembled band; ok = inside the block band
    const long hi = gi > gj ? gi : gj, d = gi > gj ? gi - gj : gj - gi;
    ok = d < band_w;
    return Aband + hi * band_w + (band_w - 1 - (ok ? d : 0));
  };
  // ---- 1. state of the block ----
  // Non-first levels, TMA path (when the staging area fits beside the working set): everything the block needs is FOUR contiguous pieces of
  // the previous level's output — rows 30..59 of the left update slot, rows 0..29 of the right one, the block's border rows and its diagonal
  // block — fetched by four bulk asynchronous copies (cp.async.bulk + mbarrier: no registers, no per-element address arithmetic) and
  // combined from shared memory. The slot rows carry the same column structure [E | F | border] as the working tile X.
  const bool tma = !kFirst && use_tma != 0;
  double* stA = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(dinv + 32) + (size_t(nbw + 1) / 2) * 8 + 15) & ~uintptr_t(15));
  double* stB = stA + size_t(kCrB) * UW;
  double* stC = stB + size_t(kCrB) * UW;
  double* stD = stC + size_t(kCrB) * nbw;
  __shared__ __align__(8) unsigned long long s_bar;
  if (tma) {
    if (t == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (t == 0) {
      const unsigned bU = unsigned(sizeof(double) * kCrB * UW), bB = unsigned(sizeof(double) * kCrB * nbw), bD = unsigned(sizeof(double) * kCrB * kCrB);
      mbar_expect_tx(&s_bar, (UL ? bU : 0u) + (UR ? bU : 0u) + bB + bD);
      if (UL) bulk_g2s(stA, UL + size_t(kCrB) * UW, bU, &s_bar);
      if (UR) bulk_g2s(stB, UR, bU, &s_bar);
      bulk_g2s(stC, crBd, bB, &s_bar);
      bulk_g2s(stD, crD, bD, &s_bar);
    }
    mbar_wait(&s_bar, 0);
    for (int e = t; e < kCrB * kCrB; e += kCrThreads) {
      const int r = e / kCrB, c = e - r * kCrB;
      Dm[r * kCrLs + c] = stD[e] - (UL ? stA[size_t(r) * UW + kCrB + c] : 0.0) - (UR ? stB[size_t(r) * UW + c] : 0.0);
    }
    for (int e = t; e < kCrB * nbw; e += kCrThreads) {
      const int r = e / nbw, c = e - r * nbw;
      X[r * XS + 2 * kCrB + c] = stC[e] - (UL ? stA[size_t(r) * UW + 2 * kCrB + c] : 0.0) - (UR ? stB[size_t(r) * UW + 2 * kCrB + c] : 0.0);
    }
    if (elim) {
      // couplings to the active neighbours: E = minus the E-part of the left slot's rows, F = minus the F-part of the right slot's rows
      // (the 60 x 60 corner of a slot is symmetric: U[r][30 + c] == U[30 + c][r] bit for bit)
      for (int e = t; e < kCrB * 2 * kCrB; e += kCrThreads) {
        const int r = e / (2 * kCrB), c = e - r * (2 * kCrB);
        const bool left = c < kCrB;
        const double v = left ? ((has_a && UL) ? -stA[size_t(r) * UW + c] : 0.0) : ((has_b && UR) ? -stB[size_t(r) * UW + c] : 0.0);
        X[r * XS + c] = v;
      }
    }
  } else {
  cr_batched(kCrB * kCrB, t, [&](int e) -> double {
    const int r = e / kCrB, c = e - r * kCrB;
    if (kFirst) {
      const bool in = r0 + r < n && r0 + c < n;
      const long gi = g0 + min(r0 + r, n - 1), gj = g0 + min(r0 + c, n - 1);
      bool ok;
      const double a = *band_ptr(gi, gj, ok), dd = dtil2 ? dtil2[gi] : 0.0;
      return in ? a + (r == c ? dd : 0.0) : (r == c ? 1.0 : 0.0);   // padding rows of a partial last block: identity
    }
    return crD[e] - ULz[size_t(kCrB + r) * UW + kCrB + c] - URz[size_t(r) * UW + c];
  }, [&](int e, double v) { const int r = e / kCrB; Dm[r * kCrLs + (e - r * kCrB)] = v; });
  if (kFirst) {
    // border = [separator columns (cal0 of them, only with several chunks) | calibration columns: rows of B | rhs: gradient]
    const int cal0 = sy.cal0, ncal = nbw - 1 - cal0;
    cr_batched(kCrB * ncal, t, [&](int e) -> double {
      const int r = e / ncal, c = e - r * ncal;
      const double v = Bmat[(g0 + min(r0 + r, n - 1)) * b_stride + c];
      return r0 + r < n ? v : 0.0;
    }, [&](int e, double v) { const int r = e / ncal; X[r * XS + 2 * kCrB + cal0 + (e - r * ncal)] = v; });
    if (t < kCrB) X[t * XS + 2 * kCrB + nbw - 1] = r0 + t < n ? grad[(g0 + min(r0 + t, n - 1)) * rhs_stride] : 0.0;
    if (cal0 > 0) cr_batched(kCrB * cal0, t, [&](int e) -> double {
      const int r = e / cal0, c = e - r * cal0;
      const long gi = g0 + min(r0 + r, n - 1), gj = colg[c];
      bool ok;
      const double v = *band_ptr(gi, gj < 0 ? gi : gj, ok);
      return (r0 + r < n && ok && gj >= 0) ? v : 0.0;
    }, [&](int e, double v) { const int r = e / cal0; X[r * XS + 2 * kCrB + (e - r * cal0)] = v; });
  } else {
    cr_batched(kCrB * nbw, t, [&](int e) -> double {
      return crBd[e] - ULz[size_t(kCrB + e / nbw) * UW + 2 * kCrB + e % nbw] - URz[size_t(e / nbw) * UW + 2 * kCrB + e % nbw];
    }, [&](int e, double v) { const int r = e / nbw; X[r * XS + 2 * kCrB + (e - r * nbw)] = v; });
  }
  }
  if (!elim) {
    if (zid != 0) return;                                        // a surviving block only carries its state over: one CTA does it
    __syncthreads();
    for (int e = t; e < kCrB * kCrB; e += kCrThreads) sy.crD[size_t(i) * (kCrB * kCrB) + e] = Dm[(e / kCrB) * kCrLs + e % kCrB];
    for (int e = t; e < kCrB * nbw; e += kCrThreads) sy.crBd[size_t(i) * kCrB * nbw + e] = X[(e / nbw) * XS + 2 * kCrB + e % nbw];
    return;
  }
  // couplings to the active neighbours: E = A(i, a) (rows of i, columns of a), F = A(i, b)
  if (!tma) cr_batched(kCrB * 2 * kCrB, t, [&](int e) -> double {
    const int r = e / (2 * kCrB), c = e - r * (2 * kCrB);
    const bool left = c < kCrB;
    const int cb = left ? c : c - kCrB;
    if (kFirst) {
      const int col = left ? r0 - kCrB + cb : r0 + kCrB + cb;       // chunk row of the neighbour's unknown
      const bool in = r0 + r < n && col >= 0 && col < n && (left ? has_a : has_b);
      bool ok;
      const double v = *band_ptr(g0 + min(r0 + r, n - 1), g0 + min(max(col, 0), n - 1), ok);
      return (in && ok) ? v : 0.0;
    }
    const double v = left ? ULz[size_t(kCrB + r) * UW + cb] : URz[size_t(kCrB + cb) * UW + r];
    return (left ? has_a : has_b) ? -v : 0.0;
  }, [&](int e, double v) { const int r = e / (2 * kCrB); X[r * XS + (e - r * (2 * kCrB))] = v; });
  for (int e = t; e < 2 * XS; e += kCrThreads) X[kCrB * XS + e] = 0.0;   // k-padding rows 30, 31
  __syncthreads();
  CB2_CLK(1);
  // ---- 2a. Cholesky of the diagonal block, 6 columns per step, with a one-step LOOKAHEAD: the one-thread factorisation of the 6x6 pivot
  //      block — the only serial piece — is taken off the critical path: while warps 1..7 apply step s to the rest of the trailing matrix,
  //      warp 0 updates the NEXT pivot block first and factors it. Two block barriers per step instead of three. ----
  auto factor_pivot = [&](int c0) {                              // one thread
    double a[6][6];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c <= r; ++c) a[r][c] = Dm[(c0 + r) * kCrLs + c0 + c];
    int fail = 0;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      double d = a[c][c];
      if (!(d > 0.0) || !isfinite(d)) { fail = 1; d = 1.0; }
      const double inv = rsqrt(d);
      a[c][c] = d * inv;
      dinv[c0 + c] = inv;
#pragma unroll
      for (int r = c + 1; r < 6; ++r) a[r][c] *= inv;
#pragma unroll
      for (int r = c + 1; r < 6; ++r)
#pragma unroll
        for (int k = c + 1; k <= r; ++k) a[r][k] -= a[r][c] * a[k][c];
    }
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c <= r; ++c) Dm[(c0 + r) * kCrLs + c0 + c] = a[r][c];
    if (fail) s_fail = 1;
  };
  if (t == 0) factor_pivot(0);
  __syncthreads();
  for (int c0 = 0; c0 < kCrB; c0 += 6) {
    const int rem = kCrB - c0 - 6;                               // rows below the pivot block
    if (t < rem) {                                               // panel row: forward-solve against the pivot block
      double* pr = Dm + (c0 + 6 + t) * kCrLs + c0;
      double x[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double s = pr[c];
#pragma unroll
        for (int k = 0; k < c; ++k) s -= x[k] * Dm[(c0 + c) * kCrLs + c0 + k];
        x[c] = s * dinv[c0 + c];
      }
#pragma unroll
      for (int c = 0; c < 6; ++c) pr[c] = x[c];
    }
    __syncthreads();
    if (rem > 0) {
      if (t < 32) {                                              // warp 0: the next pivot block, then its factorisation
        if (t < 21) {
          int a6 = 0, r6 = t; while (r6 > a6) { r6 -= a6 + 1; ++a6; }
          const int b6 = r6;                                     // entry (a6, b6), b6 <= a6
          const double* lr = Dm + (c0 + 6 + a6) * kCrLs + c0;
          const double* lc = Dm + (c0 + 6 + b6) * kCrLs + c0;
          double sacc = 0.0;
#pragma unroll
          for (int k = 0; k < 6; ++k) sacc += lr[k] * lc[k];
          Dm[(c0 + 6 + a6) * kCrLs + c0 + 6 + b6] -= sacc;
        }
        __syncwarp();
        if (t == 0) factor_pivot(c0 + 6);
      } else {
        for (int e = t - 32; e < rem * rem; e += kCrThreads - 32) {   // rest of the trailing update (lower triangle, rows beyond the next pivot block)
          const int rr = e / rem, cc = e % rem;
          if (cc > rr || rr < 6) continue;
          const double* lr = Dm + (c0 + 6 + rr) * kCrLs + c0;
          const double* lc = Dm + (c0 + 6 + cc) * kCrLs + c0;
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < 6; ++k) s += lr[k] * lc[k];
          Dm[(c0 + 6 + rr) * kCrLs + c0 + 6 + cc] -= s;
        }
      }
    }
    __syncthreads();
  }
  for (int e = t; e < kCrB * kCrLc; e += kCrThreads) { const int r = e / kCrLc, c = e % kCrLc; Lc[e] = (c < r) ? Dm[r * kCrLs + c] : 0.0; }
  __syncthreads();
  CB2_CLK(2);
  // ---- 2b. W = L^-1 X for the coupling columns and the own slice, one thread per column (L is read as a shared-memory broadcast) ----
  for (int idx = t; idx < 2 * kCrB + n_extra; idx += kCrThreads) {
    const int c = idx < 2 * kCrB ? idx : sl_lo + (idx - 2 * kCrB);
    double w[kCrB];
#pragma unroll
    for (int r = 0; r < kCrB; ++r) {
      // four interleaved partial sums shorten the dependent FMA chain; the factor row comes in 16-byte broadcast loads
      double s0 = X[r * XS + c], s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
      for (int k = 0; k + 1 < r; k += 2) {
        const double2 l = *reinterpret_cast<const double2*>(Lc + r * kCrLc + k);
        if ((k & 2) == 0) { s0 -= l.x * w[k]; s1 -= l.y * w[k + 1]; } else { s2 -= l.x * w[k]; s3 -= l.y * w[k + 1]; }
      }
      if (r & 1) s0 -= Lc[r * kCrLc + r - 1] * w[r - 1];
      w[r] = ((s0 + s1) + (s2 + s3)) * dinv[r];
    }
#pragma unroll
    for (int r = 0; r < kCrB; ++r) X[r * XS + c] = w[r];
    if (c < 2 * kCrB) {
      if (zid == 0) {
#pragma unroll
        for (int r = 0; r < kCrB; ++r) sy.crWef[size_t(i) * kCrB * (2 * kCrB) + c * kCrB + r] = w[r];   // transposed: [60][30]
      }
    } else {
#pragma unroll
      for (int r = 0; r < kCrB; ++r) if (r0 + r < n) sy.W[size_t(r0 + r) * nbw + (c - 2 * kCrB)] = w[r];
    }
  }
  if (zid == 0) for (int e = t; e < kCrB * kCrB; e += kCrThreads) {
    const int r = e / kCrB, c = e % kCrB;
    sy.crL[size_t(i) * (kCrB * kCrB) + e] = c <= r ? Dm[r * kCrLs + c] : 0.0;
  }
  __syncthreads();
  CB2_CLK(3);
  if (t == 0 && s_fail && zid == 0) atomicAdd(&scal[kScSolveFail], 1.0);
  if (nact == 1) return;                                         // last block of the chunk: nothing left to update
  // ---- 2c. U = [W_E W_F]^T W on the FP64 tensor pipe: 8x8 tiles, k = 32 (rows 30, 31 are zero) ----
  {
    double* U = sy.crU + size_t(level & 1) * sy.cr_uslots * usz + size_t(i / (2 * stride)) * usz;
    const int warp = t >> 5, lane = t & 31, fr = lane & 3, fc = lane >> 2;
    const int nqb = qb_hi;                                       // own slice of column blocks [qb_lo, qb_hi)
    // Warp w owns the 8 output rows p = 8 w .. 8 w + 7 (60 rows = 7.5 row blocks = the 8 warps): its A fragments (8 k-steps) stay in
    // registers, the B fragments stream from shared memory, 4 column blocks (independent accumulators) at a time.
    const int pb = warp;
    double af[8];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) af[ks] = X[(4 * ks + fr) * XS + 8 * pb + fc];
    const int p = 8 * pb + fc;
    for (int qb0 = qb_lo; qb0 < nqb; qb0 += 4) {
      double acc[4][2];
#pragma unroll
      for (int u = 0; u < 4; ++u) { acc[u][0] = 0.0; acc[u][1] = 0.0; }
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int qb = min(qb0 + u, nqb - 1);
          dmma_8x8x4(acc[u][0], acc[u][1], af[ks], X[(4 * ks + fr) * XS + 8 * qb + fc]);
        }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int q = 8 * (qb0 + u) + 2 * fr;
        if (qb0 + u < nqb && p < 2 * kCrB) {
          if (q < M) U[size_t(p) * UW + q] = acc[u][0];
          if (q + 1 < M) U[size_t(p) * UW + q + 1] = acc[u][1];
        }
      }
    }
  }
#if defined(CB2_CR_CLOCKS) && !defined(CB2_EMUL)
  __syncthreads();
  CB2_CLK(4);
  if (t == 0 && blockIdx.x == 1 && (level == 0 || level == 2))
    printf("[cr clocks] level %d: load %lld chol %lld trisolve+store %lld dmma %lld total %lld\n", level, clk_[1] - clk_[0], clk_[2] - clk_[1], clk_[3] - clk_[2],
           clk_[4] - clk_[3], clk_[4] - clk_[0]);
#endif
}

// Back-substitution of the levels level_hi .. level_lo (descending): x_i = L^-T (v_i - W_E x_a - W_F x_b) for the blocks eliminated
// on each level; one warp per block. grid = (ceil(eliminated blocks / 8), chunks), block = 256. Several levels per launch only
// when every one of them fits one CTA (<= 8 eliminated blocks): the levels are then separated by a block barrier.
// Everything a block needs is fetched with independent, coalesced loads up front (factor rows and [W_E | W_F]^T in registers), so the
// 30 dependent steps of the triangular solve cost one shuffle + one FMA each.
// grid_sync != nullptr: ALL levels level_hi .. level_lo in one launch whatever their size — the CTAs (all co-resident: the host launches
// at most one per SM) meet at a global barrier between levels (an arrival counter in HBM, zeroed before the launch), which removes one
// kernel launch + drain per level from the dependent chain; ytil is then read around L1 (other SMs wrote it during this launch).
// cluster != 0: the launch is ONE thread-block cluster per chunk (<= 8 CTAs = 64 eliminated blocks per level): the levels are separated by
// the hardware cluster barrier (barrier.cluster, release / acquire) — C4: the levels with <= 64 eliminated blocks (8 of 10) in one launch.
__global__ void __launch_bounds__(256) cr_back_kernel(const BandSys* __restrict__ systems, int level_hi, int level_lo, double* ytil,
                                                      unsigned* grid_sync = nullptr, int cluster = 0) {
  const BandSys sy = systems[blockIdx.y];
  unsigned sync_round = 0;
  auto yld = [&](int idx) -> double {
#if defined(CB2_EMUL)
    return ytil[idx];
#else
    return (grid_sync || cluster) ? __ldcg(ytil + idx) : ytil[idx];
#endif
  };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = sy.n;
  for (int level = level_hi; level >= level_lo; --level) {
    if (level < cr_levels(sy.nblk)) {
      const int stride = 1 << level;
      const int nact = (sy.nblk + stride - 1) >> level;
      const int q = blockIdx.x * 8 + warp;                           // q-th eliminated block of this level
      const int j = nact == 1 ? 0 : 2 * q + 1;
      if (j < nact && !(nact == 1 && q > 0)) {
        const int i = j * stride, r0 = i * kCrB;
        const bool has_a = (j & 1) != 0, has_b = (j & 1) && i + stride < sy.nblk;
        const int ra = (i - stride) * kCrB, rb = (i + stride) * kCrB;
        const int ll = lane < kCrB ? lane : kCrB - 1;
        const double* __restrict__ L = sy.crL + size_t(i) * (kCrB * kCrB);
        const double* __restrict__ WefT = sy.crWef + size_t(i) * kCrB * (2 * kCrB);   // [60][30]: column c of [W_E | W_F], contiguous over rows
        double Lr[kCrB];
#pragma unroll
        for (int c = 0; c < kCrB; ++c) Lr[c] = L[c * kCrB + ll];      // lane l holds L[c][l] for every c (row c of L = column c of L^T)
        // neighbour solutions: lane l holds x_a[l] and x_b[l]
        double xa = 0.0, xb = 0.0, rhs = 0.0;
        if (lane < kCrB) {
          if (has_a && ra + lane < n) xa = yld(sy.row_gidx[ra + lane]);
          if (has_b && rb + lane < n) xb = yld(sy.row_gidx[rb + lane]);
          if (r0 + lane < n) rhs = yld(sy.row_gidx[r0 + lane]);
        }
        if (has_a || has_b) {
          double wa[kCrB], wb[kCrB];
#pragma unroll
          for (int c = 0; c < kCrB; ++c) { wa[c] = WefT[c * kCrB + ll]; wb[c] = WefT[(kCrB + c) * kCrB + ll]; }
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int c = 0; c < kCrB; ++c) {
            s0 += wa[c] * __shfl_sync(0xffffffffu, xa, c);
            s1 += wb[c] * __shfl_sync(0xffffffffu, xb, c);
          }
          rhs -= s0 + s1;
        }
        // L^T x = rhs, backwards; lane k owns component k. Reciprocal diagonals are computed by all lanes at once.
        double dsel = 1.0;
#pragma unroll
        for (int c = 0; c < kCrB; ++c) if (lane == c) dsel = Lr[c];
        const double dinv = 1.0 / dsel;
        double x = 0.0;
#pragma unroll
        for (int c = kCrB - 1; c >= 0; --c) {
          const double xc = __shfl_sync(0xffffffffu, rhs * dinv, c);
          if (lane == c) x = xc;
          if (lane < c) rhs -= Lr[c] * xc;
        }
        if (lane < kCrB && r0 + lane < n) ytil[sy.row_gidx[r0 + lane]] = x;
      }
    }
    if (level > level_lo) {
      if (cluster) {
#if !defined(CB2_EMUL)
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
#endif
      } else if (grid_sync) {
#if !defined(CB2_EMUL)
        // global barrier: every CTA of the grid arrives once per level
        __syncthreads();
        if (threadIdx.x == 0) {
          __threadfence();
          const unsigned target = (++sync_round) * gridDim.x * gridDim.y;
          atomicAdd(grid_sync, 1u);
          while (atomicAdd(grid_sync, 0u) < target) { }
          __threadfence();
        }
        __syncthreads();
#endif
      } else { __threadfence_block(); __syncthreads(); }
    }
  }
}

}  // namespace cb2
