// calico_b200 — K5/K6/K7: Schur elimination of the block-banded control-point system and the dense reduced solve.
//
// Replaces Ceres's DENSE_SCHUR linear solver (Ceres external; batch_optimizer.cpp:12) for the LM step
//     (H + D^2) y = g,   unknowns [control points | calibration].
// Ceres's automatic ordering eliminates only every k-th control point and factors a DENSE reduced system of
// ~5 n_cp + N_c unknowns. Here the band the reference notes but does not exploit
// (bspline.hpp:287-289) is used, with time substructuring so the sequential band factorisation parallelises:
//   level 1: the control points are cut into P chunks separated by k-1 = 5 "separator" control points. Chunk interiors
//            do not couple to each other; each chunk is one banded SPD system (half bandwidth 35) with a dense border
//            [left separator 30 | right separator 30 | calibration N_c | rhs 1]. One CTA per chunk factors the band
//            (right-looking, 6 columns per step, window in shared memory) and forward-substitutes the border in the
//            same sweep (W = L^-1 border); a tiled Gram kernel forms T = W^T W.
//   level 2: the separators (30 unknowns each) form a block-tridiagonal system = banded with half bandwidth 59, border
//            [calibration | rhs], built from H minus the level-1 Schur terms; same factor + Gram kernels.
//   level 3: dense Cholesky of the N_c x N_c calibration system; then back-substitution level 2, level 1.
// The same chunk boundaries are the multi-GPU shard boundaries.
#pragma once
#include "cb2_device.cuh"

namespace cb2 {

struct BandSys {
  int n, hb, nbw;           // rows (multiple of 6), half bandwidth (hb + 1 = window size), border width (last column = rhs)
  int ksplit;               // row splits of the Gram product
  int kfirst;               // first partial the consumers sum (> 0: gram_prereduce_kernel has folded the partials before it into it)
  const int* row_gidx;      // [n] global unknown index of each row
  const int* col_gidx;      // [nbw - 1] global unknown index of each border column, -1 = absent (zero column)
  double* L;                // [n][hb+1]: A on entry, Cholesky factor on exit; (i, j) at L[i*(hb+1) + hb - (i-j)]
  double* W;                // [n][nbw]: border on entry, L^-1 border on exit
  double* T;                // [ksplit][nbw][nbw] partial Gram matrices W^T W
  double* Dinv;             // [n] reciprocal diagonal of the factor (written by band_factor_kernel, used by the back-substitution)
  int cal0;                 // first calibration column of the border: 60 with separator columns, 0 for a single chunk without separators
  // block cyclic reduction (cb2_cr.cuh): state of the 30x30 blocks, nblk = ceil(n / 30)
  int nblk, cr_uslots;
  double* crD;              // [nblk][30*30] diagonal blocks of the still-active blocks
  double* crBd;             // [nblk][30][nbw] their border rows
  const double* crZero;     // one all-zero U slot: stands in for the update of a neighbour that does not exist
  double* crU;              // [2][cr_uslots][60][60+nbw] Schur updates of the blocks eliminated on the previous / current level
  double* crWef;            // [nblk][60][30] (L^-1 [E | F])^T of every eliminated block (back-substitution)
  double* crL;              // [nblk][30*30] Cholesky factors of the diagonal blocks
};

constexpr int kSepDim = 30;   // (k-1) control points * 6
constexpr int kFacThreads = 256;

// H(gi, gj) from the assembled normal equations; gi, gj are global unknown indices.
CB2_HD size_t factor_smem_bytes(int S, int nbw) { return (size_t(S) * S + size_t(S + 6) * nbw) * sizeof(double); }

CB2_D double hess_lookup(long gi, long gj, long n_a, int N_c, const double* __restrict__ Aband, const double* __restrict__ Bmat,
                         const double* __restrict__ Cmat) {
  if (gi < gj) { const long tmp = gi; gi = gj; gj = tmp; }
  if (gi < n_a) { const long d = gi - gj; return d < kCpCols ? Aband[gi * kCpCols + (kCpCols - 1 - d)] : 0.0; }
  if (gj < n_a) return Bmat[gj * N_c + (gi - n_a)];
  return Cmat[(gi - n_a) * N_c + (gj - n_a)];
}

// Fills L (band) and W (border) of every level-1 system from the normal equations, adding the LM damping dtil2 to the diagonal.
// grid = (row blocks, systems)
__global__ void __launch_bounds__(256) gather_level1_kernel(const BandSys* __restrict__ systems, long n_a, int N_c,
                                                            const double* __restrict__ Aband, const double* __restrict__ Bmat,
                                                            const double* __restrict__ Cmat, const double* __restrict__ grad,
                                                            const double* __restrict__ dtil2) {
  const BandSys sy = systems[blockIdx.y];
  const int S = sy.hb + 1;
  const long per_row = S + sy.nbw;
  const long total = long(sy.n) * per_row;
  for (long idx = long(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += long(gridDim.x) * blockDim.x) {
    const int row = int(idx / per_row), e = int(idx % per_row);
    const long gi = sy.row_gidx[row];
    if (e < S) {
      const int col = row - (sy.hb - e);
      double v = 0.0;
      if (col >= 0) {
        const long gj = sy.row_gidx[col];
        v = hess_lookup(gi, gj, n_a, N_c, Aband, Bmat, Cmat);
        if (col == row) v += dtil2[gi];
      }
      sy.L[size_t(row) * S + e] = v;
    } else {
      const int c = e - S;
      double v;
      if (c == sy.nbw - 1) v = grad[gi];
      else { const long gj = sy.col_gidx[c]; v = gj >= 0 ? hess_lookup(gi, gj, n_a, N_c, Aband, Bmat, Cmat) : 0.0; }
      sy.W[size_t(row) * sy.nbw + c] = v;
    }
  }
}

// In-place banded Cholesky (right-looking, 6 columns per step) fused with the forward substitution of the border.
// Shared memory: Wd[S][S] circular window of the band, Wr[S][nbw] ring of partially updated border rows, Xs[6][nbw].
// Requires n % 6 == 0 and the block structure described in the header (entries beyond the block band are structurally
// zero). One CTA of 8 warps per system. The 6x6 diagonal-block factorisation — the only inherently serial piece — is taken
// off the critical path by a one-step lookahead: while warps 1..7 apply step s to the trailing window and the border, warp 0
// updates the NEXT diagonal block first and factors it. Per step, two block-wide phases:
//   P2  (needs the factored diagonal block of this step) warp 0: panel rows, 6-step triangular solve, finished factor columns
//       to HBM; warps 1..7: border rows j0..j0+5, same solve per border column -> Xs and HBM;
//   P3  warp 0: next diagonal block -= panel * panel^T, then its Cholesky (right-looking, one rsqrt per column);
//       warps 1..7: rest of the trailing window; border ring update with the 6x6 factor block held in registers and reused over
//       all border columns of a lane; every thread finally drops its software-prefetched incoming rows (global loads issued at
//       the top of the step) into the slots of the retired rows.
template <int NBLK>
__global__ void __launch_bounds__(kFacThreads) band_factor_kernel(const BandSys* __restrict__ systems, double* __restrict__ scal) {
  constexpr int S = 6 * NBLK;
  constexpr int NB1 = NBLK - 1;                       // trailing window, in 6x6 blocks
  constexpr int NTB = NB1 * (NB1 + 1) / 2;            // lower-triangular 6x6 blocks of the trailing window
  constexpr int PFB = (6 * S + kFacThreads - 1) / kFacThreads;
  constexpr int PFW = 12;                             // 6 * nbw / kFacThreads <= 12 for nbw <= 512
  constexpr int NW = kFacThreads - 32;                // worker threads (warps 1..7)
  const BandSys sy = systems[blockIdx.x];
  const int n = sy.n, nbw = sy.nbw, t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  double* Wd = dyn_smem<double>();
  double* Wr = Wd + S * S;
  double* Xs = Wr + S * nbw;          // [6][nbw] finished border rows of the current step
  __shared__ __align__(16) double Ld[2][36];
  __shared__ double Linv[2][6];
  __shared__ int s_fail;
  __shared__ unsigned char blk_i[NTB], blk_j[NTB];
  if (t == 0) {
    s_fail = 0;
    int e = 0;
    for (int bi = 0; bi < NB1; ++bi) for (int bj = 0; bj <= bi; ++bj) { blk_i[e] = (unsigned char)bi; blk_j[e] = (unsigned char)bj; ++e; }
  }
  double* __restrict__ Lg = sy.L;
  double* __restrict__ Wg = sy.W;
  // Factors the 6x6 diagonal block whose window slot is `slot` (global row jg); one thread.
  auto factor_diag = [&](int slot, int jg, int buf) {
    double a[6][6];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c <= r; ++c) a[r][c] = Wd[(slot + r) * S + slot + c];
    int fail = 0;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      double d = a[c][c];
      if (!(d > 0.0) || !isfinite(d)) { fail = 1; d = 1.0; }
      const double inv = rsqrt(d);
      a[c][c] = d * inv;
      Linv[buf][c] = inv;
      sy.Dinv[jg + c] = inv;
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
      for (int c = 0; c < 6; ++c) Ld[buf][r * 6 + c] = c <= r ? a[r][c] : 0.0;
    if (fail) s_fail = 1;
  };
  // Initial window: rows 0..min(S,n)-1.
  const int n0 = min(S, n);
  for (int e = t; e < n0 * S; e += kFacThreads) {
    const int i = e / S, d = e % S;
    const int j = i - (S - 1 - d);
    if (j >= 0) Wd[i * S + j] = Lg[size_t(i) * S + d];
  }
  for (int e = t; e < n0 * nbw; e += kFacThreads) Wr[e] = Wg[e];
  __syncthreads();
  if (t == 0 && n > 0) factor_diag(0, 0, 0);
  __syncthreads();
  int cur = 0;
  for (int j0 = 0; j0 < n; j0 += 6, cur ^= 1) {
    const int jm = j0 % S;                       // window slot of column/row j0 (6 consecutive slots, never wraps)
    const int r_end = min(j0 + S, n);            // rows j0+6 .. r_end-1 form the panel / trailing window
    const int i0 = j0 + S;                       // rows i0 .. i0+5 enter the window at the end of this step
    const bool incoming = i0 < n;
    const double* __restrict__ Ldc = Ld[cur];
    const double* __restrict__ Lic = Linv[cur];
    // P0: software prefetch of the incoming rows.
    double pfb[PFB], pfw[PFW];
    if (incoming) {
#pragma unroll
      for (int u = 0; u < PFB; ++u) { const int e = t + u * kFacThreads; if (e < 6 * S) pfb[u] = Lg[size_t(i0) * S + e]; }
#pragma unroll
      for (int u = 0; u < PFW; ++u) { const int e = t + u * kFacThreads; if (e < 6 * nbw) pfw[u] = Wg[size_t(i0) * nbw + e]; }
    }
    // P2.
    if (warp == 0) {
      // panel rows r: L(r, j0+c) = (A(r, j0+c) - sum_{k<c} L(r, j0+k) Ld[c][k]) / Ld[c][c]; emit to HBM.
      for (int rr = lane; rr < S - 6; rr += 32) {
        const int r = j0 + 6 + rr;
        if (r < r_end) {
          int rm = jm + 6 + rr; if (rm >= S) rm -= S;
          double* pr = Wd + rm * S + jm;
          double x[6];
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            double sacc = pr[c];
#pragma unroll
            for (int k = 0; k < c; ++k) sacc -= x[k] * Ldc[c * 6 + k];
            x[c] = sacc * Lic[c];
          }
#pragma unroll
          for (int c = 0; c < 6; ++c) { pr[c] = x[c]; Lg[size_t(r) * S + (S - 1 - (r - j0 - c))] = x[c]; }
        }
      }
      for (int e = lane; e < 36; e += 32) { const int r = e / 6, k = e % 6; if (k <= r) Lg[size_t(j0 + r) * S + (S - 1 - (r - k))] = Ldc[e]; }
    } else {
      // border rows j0..j0+5: forward-solve with the diagonal block -> Xs and HBM.
      for (int c = t - 32; c < nbw; c += NW) {
        double x[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          double sacc = Wr[(jm + k) * nbw + c];
#pragma unroll
          for (int q = 0; q < k; ++q) sacc -= Ldc[k * 6 + q] * x[q];
          x[k] = sacc * Lic[k];
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) { Xs[k * nbw + c] = x[k]; Wg[size_t(j0 + k) * nbw + c] = x[k]; }
      }
    }
    __syncthreads();
    // P3.
    if (warp == 0) {
      // lookahead: next diagonal block (rows/cols j0+6 .. j0+11) -= panel panel^T, then factor it for the next step.
      if (j0 + 6 < n) {
        int sm = jm + 6; if (sm >= S) sm -= S;
        if (lane < 21) {
          int a6 = 0, rem = lane; while (rem > a6) { rem -= a6 + 1; ++a6; }
          const int b6 = rem;                       // entry (a6, b6), b6 <= a6
          const double* lr = Wd + (sm + a6) * S + jm;
          const double* lc = Wd + (sm + b6) * S + jm;
          double sacc = 0.0;
#pragma unroll
          for (int k = 0; k < 6; ++k) sacc += lr[k] * lc[k];
          Wd[(sm + a6) * S + sm + b6] -= sacc;
        }
        __syncwarp();
        if (lane == 0) factor_diag(sm, j0 + 6, cur ^ 1);
      }
    } else {
      const int wt = t - 32;
      // trailing update of the band window except its first diagonal block: A(r, c) -= sum_k L(r, j0+k) L(c, j0+k).
      for (int e = 36 + wt; e < NTB * 36; e += NW) {
        const int blk = e / 36, w = e % 36;
        const int a6 = w / 6, b6 = w % 6;
        const int bi = blk_i[blk], bj = blk_j[blk];
        const int rr = 6 * bi + a6, cc = 6 * bj + b6;
        if (cc > rr || j0 + 6 + rr >= r_end) continue;
        int rm = jm + 6 + rr; if (rm >= S) rm -= S;
        int cm = jm + 6 + cc; if (cm >= S) cm -= S;
        const double* lr = Wd + rm * S + jm;
        const double* lc = Wd + cm * S + jm;
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) sacc += lr[k] * lc[k];
        Wd[rm * S + cm] -= sacc;
      }
      // border ring update, one 6-row batch per warp trip: Wr(r, c) -= sum_k L(r, j0+k) x_k(c); the 6x6 factor block stays
      // in registers and is reused for every border column of the lane.
      const int nbatch = (r_end - (j0 + 6)) / 6;
      for (int bq = warp - 1; bq < nbatch; bq += kFacThreads / 32 - 1) {
        int rm = jm + 6 + 6 * bq; if (rm >= S) rm -= S;
        double lb[6][6];
#pragma unroll
        for (int v = 0; v < 6; ++v) {
          const double2* lr = reinterpret_cast<const double2*>(Wd + (rm + v) * S + jm);
          const double2 l0 = lr[0], l1 = lr[1], l2 = lr[2];
          lb[v][0] = l0.x; lb[v][1] = l0.y; lb[v][2] = l1.x; lb[v][3] = l1.y; lb[v][4] = l2.x; lb[v][5] = l2.y;
        }
        for (int c = lane; c < nbw; c += 32) {
          double x[6], acc[6];
#pragma unroll
          for (int k = 0; k < 6; ++k) x[k] = Xs[k * nbw + c];
#pragma unroll
          for (int v = 0; v < 6; ++v) acc[v] = Wr[(rm + v) * nbw + c];
#pragma unroll
          for (int v = 0; v < 6; ++v)
#pragma unroll
            for (int k = 0; k < 6; ++k) acc[v] -= lb[v][k] * x[k];
#pragma unroll
          for (int v = 0; v < 6; ++v) Wr[(rm + v) * nbw + c] = acc[v];
        }
      }
    }
    // P4: the prefetched rows take the slots of the retired rows j0..j0+5 (nobody reads those slots in P3).
    if (incoming) {
#pragma unroll
      for (int u = 0; u < PFB; ++u) {
        const int e = t + u * kFacThreads;
        if (e < 6 * S) {
          const int i = i0 + e / S, d = e % S;
          const int j = i - (S - 1 - d);
          Wd[(i % S) * S + (j % S)] = pfb[u];
        }
      }
#pragma unroll
      for (int u = 0; u < PFW; ++u) {
        const int e = t + u * kFacThreads;
        if (e < 6 * nbw) Wr[jm * nbw + e] = pfw[u];      // rows i0..i0+5 land in slots jm..jm+5, contiguous
      }
    }
    __syncthreads();
  }
  if (t == 0 && s_fail) atomicAdd(&scal[kScSolveFail], 1.0);
}

// T[sys][k] = W[rows of split k]^T W[rows of split k]  (nbw x nbw, full symmetric storage).
// grid = (lower tile pairs of 64x64 tiles, systems, ksplit); block = (16, 16), 4x4 outputs per thread.
__global__ void __launch_bounds__(256) border_gram_kernel(const BandSys* __restrict__ systems) {
  __shared__ double As[16][65];
  __shared__ double Bs[16][65];
  const BandSys sy = systems[blockIdx.y];
  const int nbw = sy.nbw;
  const int nt = (nbw + 63) / 64;
  if (int(blockIdx.z) >= sy.ksplit) return;
  // decode the lower-triangular tile pair
  int ta = 0, tb = 0;
  { int rem = blockIdx.x; while (rem > ta) { rem -= ta + 1; ++ta; } tb = rem; }
  if (ta >= nt) return;
  const int rows_per = ((sy.n + sy.ksplit - 1) / sy.ksplit + 15) / 16 * 16;
  const int r_begin = blockIdx.z * rows_per, r_end = min(sy.n, r_begin + rows_per);
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 16 + tx;
  double acc[4][4];
  for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  for (int r0 = r_begin; r0 < r_end; r0 += 16) {
    __syncthreads();
    for (int e = tid; e < 16 * 64; e += 256) {
      const int rr = e / 64, cc = e % 64;
      const int r = r0 + rr;
      const int ca = ta * 64 + cc, cb = tb * 64 + cc;
      As[rr][cc] = (r < r_end && ca < nbw) ? sy.W[size_t(r) * nbw + ca] : 0.0;
      Bs[rr][cc] = (r < r_end && cb < nbw) ? sy.W[size_t(r) * nbw + cb] : 0.0;
    }
    __syncthreads();
    for (int rr = 0; rr < 16; ++rr) {
      double a4[4], b4[4];
      for (int a = 0; a < 4; ++a) { a4[a] = As[rr][ty * 4 + a]; b4[a] = Bs[rr][tx * 4 + a]; }
      for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) acc[a][b] += a4[a] * b4[b];
    }
  }
  double* T = sy.T + size_t(blockIdx.z) * nbw * nbw;
  for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) {
    const int ia = ta * 64 + ty * 4 + a, ib = tb * 64 + tx * 4 + b;
    if (ia < nbw && ib < nbw) {
      T[size_t(ia) * nbw + ib] = acc[a][b];
      if (ta != tb) T[size_t(ib) * nbw + ia] = acc[a][b];
    }
  }
}

// The same Gram product on the FP64 tensor pipe (mma.sync.m8n8k4.f64, SASS DMMA) for borders of up to kGramMaxNbw columns:
// grid = (ksplit, systems), one CTA per row split computes the whole lower block triangle (8x8 tiles, dealt round-robin to the 8
// warps, accumulators in registers) while W streams through shared memory 32 rows at a time (row stride == 4 mod 16: conflict-free
// fragment loads). A and B fragments of a Gram product are the same data: lane l holds W[4 ks + l % 4][8 b + l / 4].
constexpr int kGramRows = 32;
constexpr int kGramMaxTiles = 44;                      // per warp: 8 * 44 = 352 >= 26 * 27 / 2
constexpr int kGramMaxNbw = 208;
CB2_HD int gram_stride(int nbw) { const int w = (nbw + 7) / 8 * 8; return (w - 4 + 15) / 16 * 16 + 4; }
CB2_HD size_t gram_smem_bytes(int nbw) { return size_t(kGramRows) * gram_stride(nbw) * sizeof(double); }
CB2_D void gram_dmma(double& c0, double& c1, double a, double b) {
#if defined(CB2_EMUL)
  ::cb2emul::dmma_8x8x4(c0, c1, a, b);
#else
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
#endif
}
// Row subsets (cyclic reduction only): blk_mod > 1 restricts the product to the 30-row blocks i that are multiples of blk_mod = 2^L
// (blk_res = 0: the blocks the reduction levels >= L eliminate) or to all the others (blk_res = -1: the blocks of the levels < L, whose W
// rows are final once level L - 1 has run): the host launches the second kind beside the later, narrow, latency-bound levels. A launch
// owns the partial matrices k_off .. k_off + k_cnt - 1 of T (k_cnt < 0: all sy.ksplit of them, every row).
constexpr int kGramBlk = 30;
__global__ void __launch_bounds__(256) border_gram_dmma_kernel(const BandSys* __restrict__ systems, int blk_res = 0, int blk_mod = 1, int k_off = 0,
                                                               int k_cnt = -1) {
  const BandSys sy = systems[blockIdx.y];
  const int kc = k_cnt < 0 ? sy.ksplit : k_cnt;
  if (int(blockIdx.x) >= kc) return;
  const int nbw = sy.nbw, nb = (nbw + 7) / 8, XS = gram_stride(nbw), ntile = nb * (nb + 1) / 2;
  double* tile = dyn_smem<double>();
  __shared__ unsigned char tbi[8 * kGramMaxTiles], tbj[8 * kGramMaxTiles];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31, fr = lane & 3, fc = lane >> 2;
  for (int e = t; e < ntile; e += 256) { int bi = 0, rem = e; while (rem > bi) { rem -= bi + 1; ++bi; } tbi[e] = (unsigned char)bi; tbj[e] = (unsigned char)rem; }
  const int nblk30 = (sy.n + kGramBlk - 1) / kGramBlk;
  const int nmult = (nblk30 + blk_mod - 1) / blk_mod;          // multiples of blk_mod below nblk30
  const int nsel = blk_res < 0 ? nblk30 - nmult : nmult;
  const int nv = blk_mod == 1 ? sy.n : nsel * kGramBlk;      // rows of this launch, numbered consecutively ("virtual" rows)
  auto row_of = [&](int v) {
    if (blk_mod == 1) return v;
    const int mb = v / kGramBlk;
    const int blk = blk_res < 0 ? (mb / (blk_mod - 1)) * blk_mod + mb % (blk_mod - 1) + 1 : mb * blk_mod;
    return blk * kGramBlk + (v - mb * kGramBlk);
  };
  const int rows_per = ((nv + kc - 1) / kc + kGramRows - 1) / kGramRows * kGramRows;
  const int r_begin = blockIdx.x * rows_per, r_end = min(nv, r_begin + rows_per);
  double acc[kGramMaxTiles][2];
#pragma unroll
  for (int i = 0; i < kGramMaxTiles; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
  const double* __restrict__ Wg = sy.W;
  const int wcols = nb * 8;
  for (int r0 = r_begin; r0 < r_end; r0 += kGramRows) {
    __syncthreads();
    // 6 independent loads in flight per thread (clamped addresses, selects afterwards): the chunk load is pure latency otherwise.
    for (int e0 = t; e0 < kGramRows * wcols; e0 += 6 * 256) {
      double v[6];
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int e = min(e0 + u * 256, kGramRows * wcols - 1);
        const int rr = e / wcols, cc = e - rr * wcols;
        const int ar = row_of(min(r0 + rr, nv - 1));
        const bool ok = r0 + rr < r_end && ar < sy.n && cc < nbw;
        const double x = Wg[size_t(ok ? ar : 0) * nbw + (ok ? cc : 0)];
        v[u] = ok ? x : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int e = e0 + u * 256;
        if (e < kGramRows * wcols) { const int rr = e / wcols; tile[rr * XS + (e - rr * wcols)] = v[u]; }
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int ks = 0; ks < kGramRows / 4; ++ks) {
      const double* trow = tile + (4 * ks + fr) * XS + fc;
#pragma unroll
      for (int i = 0; i < kGramMaxTiles; ++i) {
        const int e = warp + 8 * i;
        if (e < ntile) gram_dmma(acc[i][0], acc[i][1], trow[8 * tbi[e]], trow[8 * tbj[e]]);
      }
    }
  }
  double* T = sy.T + size_t(k_off + blockIdx.x) * nbw * nbw;
#pragma unroll
  for (int i = 0; i < kGramMaxTiles; ++i) {
    const int e = warp + 8 * i;
    if (e < ntile) {
      const int bi = tbi[e], bj = tbj[e];
      const int a = 8 * bi + fc;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int b = 8 * bj + 2 * fr + q;
        if (a < nbw && b < nbw) {
          T[size_t(a) * nbw + b] = acc[i][q];
          if (bi != bj) T[size_t(b) * nbw + a] = acc[i][q];
        }
      }
    }
  }
}

// Folds the partial Gram matrices k_lo .. k_hi - 1 into partial k_hi - 1 (fixed order): run beside the late reduction levels on the partials
// of the early launch, so that the consumers on the critical path (level3_build_kernel) sum 1 + k_last partials instead of ~115.
__global__ void __launch_bounds__(256) gram_prereduce_kernel(const BandSys* __restrict__ systems, int k_lo, int k_hi) {
  const BandSys sy = systems[blockIdx.y];
  const size_t stride = size_t(sy.nbw) * sy.nbw;
  for (size_t e = size_t(blockIdx.x) * blockDim.x + threadIdx.x; e < stride; e += size_t(gridDim.x) * blockDim.x) {
    double* __restrict__ T = sy.T + e;
    double s = 0.0;
    int k = k_lo;
    for (; k + 8 <= k_hi; k += 8) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = T[(k + u) * stride];
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; k < k_hi; ++k) s += T[k * stride];
    T[(k_hi - 1) * stride] = s;
  }
}

// Sum over the row splits of a system's Gram matrix.
CB2_D double gram_at(const BandSys& sy, int a, int b) {
  const size_t stride = size_t(sy.nbw) * sy.nbw;
  const double* __restrict__ T = sy.T + size_t(a) * sy.nbw + b;
  double s = 0.0;
  int k = sy.kfirst;
  for (; k + 8 <= sy.ksplit; k += 8) {   // 8 independent loads in flight, summed in the fixed order k = kfirst, kfirst + 1, ...
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = T[(k + u) * stride];
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
  }
  for (; k < sy.ksplit; ++k) s += T[k * stride];
  return s;
}

// Level-2 system (separators) = H restricted to the separators minus the level-1 Schur terms of the owned chunks.
// Separator s sits between chunk s (its right border, columns 30..59) and chunk s+1 (its left border, columns 0..29).
// chunk_sys[c] = index of chunk c in `chunks`, or -1 if the chunk is not owned by this rank (multi-GPU) or has no rows.
// rawdiag receives the undamped H diagonal of the separator rows (for the LM damping added after a cross-rank reduction).
__global__ void __launch_bounds__(256) level2_build_kernel(BandSys l2, const BandSys* __restrict__ chunks, const int* __restrict__ chunk_sys,
                                                           int n_chunks, long n_a, int N_c, const double* __restrict__ Aband,
                                                           const double* __restrict__ Bmat, const double* __restrict__ Cmat,
                                                           const double* __restrict__ grad, double* __restrict__ rawdiag) {
  const int S = l2.hb + 1;   // 60
  const long per_row = S + l2.nbw;
  const long total = long(l2.n) * per_row;
  for (long idx = long(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += long(gridDim.x) * blockDim.x) {
    const int row = int(idx / per_row), e = int(idx % per_row);
    const long gi = l2.row_gidx[row];
    const int sa = row / kSepDim, la = row % kSepDim;
    const int cR = chunk_sys[sa], cL = (sa + 1 < n_chunks) ? chunk_sys[sa + 1] : -1;   // chunk sa: separator is its right border
    if (e < S) {
      const int col = row - (l2.hb - e);
      double v = 0.0;
      if (col >= 0) {
        const int sb = col / kSepDim, lb = col % kSepDim;
        v = hess_lookup(gi, l2.row_gidx[col], n_a, N_c, Aband, Bmat, Cmat);
        if (col == row) rawdiag[row] = v;
        if (sb == sa) {
          if (cR >= 0) v -= gram_at(chunks[cR], kSepDim + la, kSepDim + lb);
          if (cL >= 0) v -= gram_at(chunks[cL], la, lb);
        } else if (sb == sa - 1) {
          if (cR >= 0) v -= gram_at(chunks[cR], kSepDim + la, lb);
        }
      }
      l2.L[size_t(row) * S + e] = v;
    } else {
      const int c = e - S;   // calibration column c, or rhs when c == N_c
      double v = (c == l2.nbw - 1) ? grad[gi] : hess_lookup(gi, n_a + c, n_a, N_c, Aband, Bmat, Cmat);
      if (cR >= 0) v -= gram_at(chunks[cR], kSepDim + la, 2 * kSepDim + c);
      if (cL >= 0) v -= gram_at(chunks[cL], la, 2 * kSepDim + c);
      l2.W[size_t(row) * l2.nbw + c] = v;
    }
  }
}

// Adds the LM damping to the level-2 diagonal (separate so a cross-rank sum can sit between build and damp).
__global__ void level2_damp_kernel(BandSys l2, const double* __restrict__ dtil2) {
  const int S = l2.hb + 1;
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < l2.n; row += gridDim.x * blockDim.x)
    l2.L[size_t(row) * S + (S - 1)] += dtil2[l2.row_gidx[row]];
}

// Level-3 (calibration) system, augmented with the rhs as an extra ROW so that the Cholesky sweep forward-substitutes it:
// Cw is [(N+1)][(N+1)] row-major; Cw[r][c] (r, c < N) = C - sum_chunks T1[cal, cal], Cw[N][c] = g_c - sum_chunks T1[rhs, cal].
// The level-2 term and the damping are applied in reduced_solve_kernel. rawdiag_c = undamped diag of C.
__global__ void __launch_bounds__(256) level3_build_kernel(const BandSys* __restrict__ chunks, int n_owned, long n_a, int N_c,
                                                           const double* __restrict__ Cmat, const double* __restrict__ grad,
                                                           double* __restrict__ Cw, double* __restrict__ rawdiag_c) {
  // 8 lanes per entry: lane `sub` sums the Gram partials k = sub, sub + 8, ... of every owned chunk (the Gram kernel leaves up to 148
  // row-split partials), then a fixed-order shuffle tree combines the 8 partial sums.
  const int ld = N_c + 1;
  const long total = long(ld) * ld;
  const int sub = threadIdx.x & 7;
  for (long idx0 = (long(blockIdx.x) * blockDim.x + threadIdx.x) >> 3; idx0 < ((total + 31) & ~31L); idx0 += (long(gridDim.x) * blockDim.x) >> 3) {
    const long idx = idx0 < total ? idx0 : total - 1;       // whole warps stay in the loop for the shuffles
    const int r = int(idx / ld), c = int(idx % ld);
    double part = 0.0;
    if (c < N_c) {
      for (int p = 0; p < n_owned; ++p) {
        const BandSys& sy = chunks[p];
        const size_t stride = size_t(sy.nbw) * sy.nbw;
        const double* __restrict__ T = sy.T + size_t(sy.cal0 + r) * sy.nbw + sy.cal0 + c;
        for (int k = sy.kfirst + sub; k < sy.ksplit; k += 8) part += T[k * stride];
      }
    }
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    part += __shfl_xor_sync(0xffffffffu, part, 4);
    if (sub == 0 && idx0 < total) {
      double v = 0.0;
      if (c < N_c) {
        v = r == N_c ? grad[n_a + c] : Cmat[size_t(r) * N_c + c];
        if (c == r) rawdiag_c[r] = v;
        v -= part;
      }
      Cw[idx] = v;
    }
  }
}

// Grid-wide: Cw[:N, :N] -= T2[:N, :N] (separator-level Schur term), += diag(dtil2_c); rhs row Cw[N][:] -= T2[N][:N].
__global__ void __launch_bounds__(256) level3_finalize_kernel(BandSys l2, int N, long n_a, double* __restrict__ Cw, const double* __restrict__ dtil2) {
  const int ld = N + 1;
  const long total = long(ld) * ld;
  for (long e = long(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += long(gridDim.x) * blockDim.x) {
    const int r = int(e / ld), c = int(e % ld);
    if (c >= N) continue;
    double v = Cw[e];
    if (l2.n > 0) v -= gram_at(l2, r, c);
    if (c == r) v += dtil2[n_a + r];
    Cw[e] = v;
  }
}

// Dense reduced solve (1 CTA) of the finalized system M = Cw[:N, :N], rhs row Cw[N][:]. Right-looking
// Cholesky in global (L2-resident) memory with 32-column panels staged in shared memory; the rhs row rides along and
// becomes z = L^-1 rhs; the backward solve L^T y = z prefetches the next factor row. y_c -> ytil[n_a + c].
// l2 may have n == 0 (no separators). Requires N <= kRedThreads.
constexpr int kRedThreads = 512;
constexpr int kRedPanel = 32;
__global__ void __launch_bounds__(kRedThreads) reduced_solve_kernel(int N, long n_a, double* __restrict__ Cw, double* __restrict__ ytil,
                                                                    double* __restrict__ scal) {
  double* Pn = dyn_smem<double>();    // [N + 1][kRedPanel + 1] panel
  __shared__ int s_fail;
  const int t = threadIdx.x;
  const int ld = N + 1;
  constexpr int PS = kRedPanel + 1;
  if (t == 0) s_fail = 0;
  __syncthreads();
  for (int p0 = 0; p0 < N; p0 += kRedPanel) {
    const int pw = min(kRedPanel, N - p0);
    const int nr = ld - p0;           // rows p0 .. N (includes the rhs row)
    for (int e = t; e < nr * pw; e += kRedThreads) { const int r = e / pw, c = e % pw; Pn[r * PS + c] = Cw[size_t(p0 + r) * ld + p0 + c]; }
    __syncthreads();
    for (int c = 0; c < pw; ++c) {
      if (t == 0) {
        double d = Pn[c * PS + c];
        if (!(d > 0.0) || !isfinite(d)) { s_fail = 1; d = 1.0; }
        Pn[c * PS + c] = sqrt(d);
      }
      __syncthreads();
      const double inv = 1.0 / Pn[c * PS + c];
      for (int r = c + 1 + t; r < nr; r += kRedThreads) Pn[r * PS + c] *= inv;
      __syncthreads();
      const int nc2 = pw - c - 1;
      for (int e = t; e < nc2 * nr; e += kRedThreads) {
        const int c2 = c + 1 + e / nr, r = e % nr;
        if (r >= c2) Pn[r * PS + c2] -= Pn[r * PS + c] * Pn[c2 * PS + c];
      }
      __syncthreads();
    }
    for (int e = t; e < nr * pw; e += kRedThreads) { const int r = e / pw, c = e % pw; Cw[size_t(p0 + r) * ld + p0 + c] = Pn[r * PS + c]; }
    // trailing update: Cw(r, c2) -= sum_k Pn(r, k) Pn(c2, k) for r >= c2 >= p0 + pw (c2 < N)
    const int ntr = nr - pw;
    for (long e = t; e < long(ntr) * ntr; e += kRedThreads) {
      const int rr = int(e / ntr), cc = int(e % ntr);
      if (cc > rr || p0 + pw + cc >= N) continue;
      const double* pr = Pn + (pw + rr) * PS;
      const double* pc = Pn + (pw + cc) * PS;
      double s = 0.0;
      for (int k = 0; k < pw; ++k) s += pr[k] * pc[k];
      Cw[size_t(p0 + pw + rr) * ld + p0 + pw + cc] -= s;
    }
    __syncthreads();
  }
  double* y = Pn;
  if (t < N) y[t] = Cw[size_t(N) * ld + t];
  __syncthreads();
  double lj_next = (N > 0 && t <= N - 1) ? Cw[size_t(N - 1) * ld + t] : 0.0;
  for (int j = N - 1; j >= 0; --j) {
    const double lj = lj_next;
    if (j > 0 && t <= j - 1) lj_next = Cw[size_t(j - 1) * ld + t];
    if (t == j) y[j] = y[j] / lj;
    __syncthreads();
    if (t < j) y[t] -= lj * y[j];
    __syncthreads();
  }
  if (t < N) ytil[n_a + t] = y[t];
  if (t == 0 && s_fail) atomicAdd(&scal[kScSolveFail], 1.0);
}

// Same solve with the whole augmented matrix resident in shared memory (N <= ~165: (N+1) x ld doubles <= 227 KB): blocked
// right-looking Cholesky, 8 columns per step: (1) thread 0 factors the 8x8 diagonal block in registers, (2) one thread per row
// solves the panel against it and drops the result both in place and into a transposed panel buffer PT[k][row], (3) all threads
// apply the rank-8 update to the trailing lower triangle in 4x4 register tiles (16-byte shared loads). The rhs row rides along.
// The backward solve runs on one warp (32-column blocks, shuffle-broadcast inside a block).
constexpr int kRsPanel = 8;
CB2_HD int rs_ld(int N) { int ld = N + 1; while (ld % 16 != 2) ++ld; return ld; }       // even (16-byte rows), 2-way conflicts at worst down a column
CB2_HD int rs_ldp(int N) { return (N + 1 + 3) / 4 * 4 + 4; }
CB2_HD size_t reduced_smem_bytes(int N) { return (size_t(N + 1) * rs_ld(N) + size_t(kRsPanel) * rs_ldp(N) + N + 1) * sizeof(double); }
__global__ void __launch_bounds__(kRedThreads) reduced_solve_smem_kernel(int N, long n_a, double* __restrict__ Cw, double* __restrict__ ytil,
                                                                         double* __restrict__ scal) {
  const int ld = rs_ld(N), ldp = rs_ldp(N), ldg = N + 1, t = threadIdx.x;
  double* A = dyn_smem<double>();            // [N + 1][ld]
  double* PT = A + size_t(N + 1) * ld;       // [8][ldp] transposed panel of the current step, indexed by row - (p0 + pw)
  double* y = PT + kRsPanel * ldp;           // [N + 1]
  __shared__ double Ld[2][kRsPanel * kRsPanel];
  __shared__ double Linv[2][kRsPanel];
  __shared__ int s_fail;
  if (t == 0) s_fail = 0;
  for (int e = t; e < ldg * ldg; e += kRedThreads) { const int r = e / ldg, c = e - r * ldg; A[r * ld + c] = Cw[e]; }
  __syncthreads();
  // Factors the pw x pw diagonal block at (p0, p0) in registers (one thread) -> Ld[buf], Linv[buf], and in place.
  auto factor_diag = [&](int p0, int pw, int buf) {
    double a[kRsPanel][kRsPanel];
#pragma unroll
    for (int r = 0; r < kRsPanel; ++r)
#pragma unroll
      for (int c = 0; c <= r; ++c) a[r][c] = (r < pw) ? A[(p0 + r) * ld + p0 + c] : (r == c ? 1.0 : 0.0);
    int fail = 0;
#pragma unroll
    for (int c = 0; c < kRsPanel; ++c) {
      double d = a[c][c];
      if (!(d > 0.0) || !isfinite(d)) { fail = 1; d = 1.0; }
      const double inv = rsqrt(d);
      a[c][c] = d * inv;
      Linv[buf][c] = inv;
#pragma unroll
      for (int r = c + 1; r < kRsPanel; ++r) a[r][c] *= inv;
#pragma unroll
      for (int r = c + 1; r < kRsPanel; ++r)
#pragma unroll
        for (int k = c + 1; k <= r; ++k) a[r][k] -= a[r][c] * a[k][c];
    }
#pragma unroll
    for (int r = 0; r < kRsPanel; ++r)
#pragma unroll
      for (int c = 0; c < kRsPanel; ++c) {
        Ld[buf][r * kRsPanel + c] = c <= r ? a[r][c] : 0.0;
        if (c <= r && r < pw) A[(p0 + r) * ld + p0 + c] = a[r][c];
      }
    if (fail) s_fail = 1;
  };
  if (t == 0 && N > 0) factor_diag(0, min(kRsPanel, N), 0);
  __syncthreads();
  int cur = 0;
  for (int p0 = 0; p0 < N; p0 += kRsPanel, cur ^= 1) {
    const int pw = min(kRsPanel, N - p0);
    const int rbase = p0 + pw;               // first row below the diagonal block
    const int nrem = ldg - rbase;            // rows rbase .. N (the last one is the rhs row)
    const double* Ldc = Ld[cur];
    const double* Lic = Linv[cur];
    for (int rr = t; rr < nrem; rr += kRedThreads) {
      double* pr = A + (rbase + rr) * ld + p0;
      double x[kRsPanel];
#pragma unroll
      for (int c = 0; c < kRsPanel; ++c) {
        double sacc = c < pw ? pr[c] : 0.0;
#pragma unroll
        for (int k = 0; k < c; ++k) sacc -= x[k] * Ldc[c * kRsPanel + k];
        x[c] = sacc * Lic[c];
      }
#pragma unroll
      for (int c = 0; c < kRsPanel; ++c) { if (c < pw) pr[c] = x[c]; PT[c * ldp + rr] = c < pw ? x[c] : 0.0; }
    }
    for (int rr = nrem + t; rr < (nrem + 3) / 4 * 4; rr += kRedThreads)      // zero the tile padding rows
#pragma unroll
      for (int c = 0; c < kRsPanel; ++c) PT[c * ldp + rr] = 0.0;
    __syncthreads();
    // trailing update: A(rbase + i, rbase + j) -= sum_k PT[k][i] PT[k][j] for i >= j, j < nrem - 1 (no rhs column).
    // Lookahead: warp 0 updates the NEXT diagonal block (i, j < 8) first and factors it while warps 1.. update everything else, so
    // the one-thread 8x8 factorisation is off the critical path and a panel step needs two block barriers instead of three.
    if (t < 32) {
      for (int e = t; e < kRsPanel * kRsPanel; e += 32) {
        const int i = e / kRsPanel, j = e - i * kRsPanel;
        if (j <= i && i < nrem && j < nrem - 1) {
          double sacc = 0.0;
#pragma unroll
          for (int k = 0; k < kRsPanel; ++k) sacc += PT[k * ldp + i] * PT[k * ldp + j];
          A[(rbase + i) * ld + rbase + j] -= sacc;
        }
      }
      __syncwarp();
      if (t == 0 && rbase < N) factor_diag(rbase, min(kRsPanel, N - rbase), cur ^ 1);
    } else {
      const int TR = (nrem + 3) / 4, TC = (nrem - 1 + 3) / 4;
      for (int e = t - 32; e < TR * TC; e += kRedThreads - 32) {
        const int tr = e / TC, tc = e - tr * TC;
        if (tc > tr || tr < 2) continue;       // tr < 2 (and hence tc < 2): the next diagonal block, owned by warp 0
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll
        for (int k = 0; k < kRsPanel; ++k) {
          const double2 r01 = *reinterpret_cast<const double2*>(PT + k * ldp + 4 * tr), r23 = *reinterpret_cast<const double2*>(PT + k * ldp + 4 * tr + 2);
          const double2 c01 = *reinterpret_cast<const double2*>(PT + k * ldp + 4 * tc), c23 = *reinterpret_cast<const double2*>(PT + k * ldp + 4 * tc + 2);
          const double rv[4] = {r01.x, r01.y, r23.x, r23.y}, cv[4] = {c01.x, c01.y, c23.x, c23.y};
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] += rv[a] * cv[b];
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int i = 4 * tr + a;
          if (i >= nrem) continue;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int j = 4 * tc + b;
            if (j <= i && j < nrem - 1) A[(rbase + i) * ld + rbase + j] -= acc[a][b];
          }
        }
      }
    }
    __syncthreads();
  }
  // backward solve L^T y = z on warp 0; z = row N of A.
  if (t < 32) {
    const int lane = t;
    for (int c = lane; c < N; c += 32) y[c] = A[N * ld + c];
    __syncwarp();
    for (int jb = (N - 1) / 32; jb >= 0; --jb) {
      const int j = 32 * jb + lane;
      const int jj = min(j, N - 1);
      double rhs = j < N ? y[j] : 0.0;
      const int k0 = 32 * (jb + 1);
      double s0 = 0.0, s1 = 0.0;
      int k = k0;
      for (; k + 1 < N; k += 2) { s0 += A[k * ld + jj] * y[k]; s1 += A[(k + 1) * ld + jj] * y[k + 1]; }
      if (k < N) s0 += A[k * ld + jj] * y[k];
      rhs -= s0 + s1;
      const double dinv = 1.0 / A[jj * ld + jj];
      double x = 0.0;
      const int cmax = min(31, N - 1 - 32 * jb);
      for (int c = cmax; c >= 0; --c) {
        const double xc = __shfl_sync(0xffffffffu, rhs * dinv, c);
        if (lane == c) x = xc;
        if (lane < c) rhs -= A[(32 * jb + c) * ld + jj] * xc;
      }
      if (j < N) y[j] = x;
      __syncwarp();
    }
    for (int c = lane; c < N; c += 32) ytil[n_a + c] = y[c];
    if (lane == 0 && s_fail) atomicAdd(&scal[kScSolveFail], 1.0);
  }
}

// Right-hand sides of the back-substitution, all rows of all systems in parallel: ytil[row] <- z - W[row, :nbw-1] . y[border].
// grid = (row blocks, systems), block = 256 (one warp per row, coalesced W reads).
__global__ void __launch_bounds__(256) border_matvec_kernel(const BandSys* __restrict__ systems, double* __restrict__ ytil) {
  const BandSys sy = systems[blockIdx.y];
  const int nbw = sy.nbw, t = threadIdx.x;
  double* coef = dyn_smem<double>();
  for (int c = t; c < nbw - 1; c += 256) { const int g = sy.col_gidx[c]; coef[c] = g >= 0 ? ytil[g] : 0.0; }
  __syncthreads();
  const int warp = t >> 5, lane = t & 31;
  for (int i = blockIdx.x * 8 + warp; i < sy.n; i += gridDim.x * 8) {
    const double* wr = sy.W + size_t(i) * nbw;
    double sacc = 0.0;
    for (int c = lane; c < nbw - 1; c += 32) sacc += wr[c] * coef[c];
    for (int off = 16; off > 0; off >>= 1) sacc += __shfl_down_sync(0xffffffffu, sacc, off);
    if (lane == 0) ytil[sy.row_gidx[i]] = wr[nbw - 1] - sacc;
  }
}

// Back-substitution of one banded system: v (from border_matvec_kernel, in ytil) -> L^T y = v (right-looking, 6 rows per step).
// grid = systems, block = 256, dynamic shared memory: 2 n doubles (v, reciprocal diagonal) + 2 * kBackRows * S
// (factor rows, double-buffered: while one batch of kBackRows rows is being used, the next one is in flight in registers).
constexpr int kBackThreads = 256;
constexpr int kBackRows = 48;
CB2_HD size_t backsolve_smem_bytes(int n, int nbw, int S) { (void)nbw; return (2 * size_t(n) + 2 * size_t(kBackRows) * S) * sizeof(double); }
__global__ void __launch_bounds__(kBackThreads) band_backsolve_kernel(const BandSys* __restrict__ systems, double* __restrict__ ytil) {
  const BandSys sy = systems[blockIdx.x];
  const int n = sy.n, S = sy.hb + 1, t = threadIdx.x;
  double* v = dyn_smem<double>();
  double* dinv = v + n;
  double* Lbuf = dinv + n;
  constexpr int PF = (kBackRows * 60 + kBackThreads - 1) / kBackThreads;   // S <= 60
  const double* __restrict__ Lg = sy.L;
  // First batch of factor rows goes straight to shared memory while the border product is computed.
  int b_end = n, b_start = max(0, n - kBackRows);
  for (int e = t; e < (b_end - b_start) * S; e += kBackThreads) Lbuf[e] = Lg[size_t(b_start) * S + e];
  for (int i = t; i < n; i += kBackThreads) { dinv[i] = sy.Dinv[i]; v[i] = ytil[sy.row_gidx[i]]; }
  __syncthreads();
  int buf = 0;
  while (b_end > 0) {
    const double* Lc = Lbuf + buf * kBackRows * S;
    const int nb_end = b_start, nb_start = max(0, b_start - kBackRows);
    double pf[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) { const int e = t + u * kBackThreads; if (e < (nb_end - nb_start) * S) pf[u] = Lg[size_t(nb_start) * S + e]; }
    for (int i0 = b_end - 6; i0 >= b_start; i0 -= 6) {
      const double* Lr = Lc + (i0 - b_start) * S;     // factor rows i0 .. i0+5
      if (t == 0) {
        double y6[6];
#pragma unroll
        for (int k = 5; k >= 0; --k) {
          double sacc = v[i0 + k];
#pragma unroll
          for (int q = k + 1; q < 6; ++q) sacc -= Lr[q * S + (S - 1 - (q - k))] * y6[q];
          y6[k] = sacc * dinv[i0 + k];
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) v[i0 + k] = y6[k];
      }
      __syncthreads();
      // v(tc) -= sum_k L(i0+k, tc) y(i0+k) for the columns tc < i0 reached by these rows
      if (t < S - 1) {
        const int tc = i0 - 1 - t;
        if (tc >= 0) {
          double sacc = 0.0;
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            const int d = k + 1 + t;           // row - col
            if (d <= S - 1) sacc += Lr[k * S + (S - 1 - d)] * v[i0 + k];
          }
          v[tc] -= sacc;
        }
      }
      __syncthreads();
    }
    double* Ln = Lbuf + (buf ^ 1) * kBackRows * S;
#pragma unroll
    for (int u = 0; u < PF; ++u) { const int e = t + u * kBackThreads; if (e < (nb_end - nb_start) * S) Ln[e] = pf[u]; }
    __syncthreads();
    b_end = nb_end; b_start = nb_start; buf ^= 1;
  }
  for (int i = t; i < n; i += kBackThreads) ytil[sy.row_gidx[i]] = v[i];
}

}  // namespace cb2
