// calico_b200 — K4: block-sparse Gauss-Newton normal equations  H = J^T J,  g = J^T r.
//
// Replaces Ceres's BlockSparseMatrix / SchurEliminator accumulation (Ceres external; selected by
// batch_optimizer.cpp:12 DENSE_SCHUR). Unknown vector = [control points 6*n_cp | calibration N_c]. Structure
// (SURVEY §8a): A = H[cp,cp] is block-banded (6x6 blocks, half-bandwidth k-1 = 5 blocks, i.e. 35 scalars) because a
// residual touches k consecutive control points (camera_cost_functor.cpp:52-60); B = H[cp,calib] couples a segment to
// the sensors observed in it; C = H[calib,calib] is block-diagonal per sensor (no residual involves two sensors).
//
// accumulate_kernel: one CTA per spline segment (per 2 / 4 segments when there are fewer sensors than warps), one WARP per
// (segment, sensor) pair, round-robin. All residual rows of a segment touch the same 36 control-point columns, so each warp forms the local Gram matrix X^T X of X = [J_cp | r | 0 0 0 | J_calib | 0..]
// (r as an extra column gives the gradient for free) on the FP64 TENSOR pipe: mma.sync.m8n8k4.f64 (SASS DMMA), 8x8
// accumulator tiles in registers, lower block triangle only. For a Gram product the A and B fragments of a column block are
// the same register (lane holds X[4 ks + lane % 4][8 b + lane / 4]), so a k-step of 4 rows costs NB shared loads for
// NB (NB + 1) / 2 DMMAs. J row tiles stream global -> shared with cp.async (per-warp double buffer, row stride 68 doubles
// == 4 mod 16: conflict-free fragment loads). The control-point block accumulates over all sensors of the warp and is
// combined across warps in a fixed order at the end; the calibration blocks are flushed per sensor. No atomics.
// Camera rows have more structure: all rows of one image share the basis weights w, and their control-point part is the Kronecker
// product g (x) w (g = d r / d pose, 6 numbers per row). For cameras with <= 16 calibration unknowns the Jacobian sweep itself (cb2_eval.cuh)
// therefore leaves per (CTA, image) the COMPACT Gram matrix of the rows [g | r | 0 | calibration] (24 x 24 instead of 56 x 56: 6 DMMA tiles
// instead of 28) while its Jacobian stores drain, and expand_gram_kernel below expands once per image:
//     H[cp_a, cp_b] += w_a w_b G_gg,   H[cp_a, r | calib] += w_a G_g.,   H[calib, calib] += G_cc.
// accumulate_kernel then only sees the IMU sensors (and cameras with more than 16 calibration unknowns): the camera Jacobian — 87 % of the
// Jacobian bytes — is written once and never read back.
// assemble_*_kernel: sums the <= 6 overlapping segment partials per control-point entry into the banded storage and
// reduces the calibration blocks over all segments.
#pragma once
#include "cb2_device.cuh"

namespace cb2 {

constexpr int kAccWarps = 4;
constexpr int kAccThreads = 32 * kAccWarps;
constexpr int kAccRows = 16;       // J rows per shared-memory tile (4 DMMA k-steps)
constexpr int kAccStride = 68;     // doubles per tile row; 68 mod 16 == 4 -> the 16 lanes of a half warp hit 16 distinct 8-byte banks
constexpr int kAccRcol = 36;       // local layout: cp 0..35 | r 36 | zeros up to CAL0 | calibration unknowns CAL0.. | zeros
constexpr int kAccCal0 = 40;       // CAL0 of the 7- and 8-block layouts; sensors with <= 11 calibration unknowns use CAL0 = 37 and 6 blocks (48 columns)
constexpr int kAccMaxSensors = 64;
CB2_HD constexpr size_t acc_smem_bytes() { return size_t(kAccWarps) * 2 * (kAccRows * kAccStride + 4) * sizeof(double); }
#ifndef CB2_ACC_MINBLOCKS
#define CB2_ACC_MINBLOCKS 3
#endif

// 8-byte asynchronous global -> shared copy (LDGSTS): the J tile of the NEXT step streams in while the current one is multiplied.
CB2_D void cp_async8(double* dst, const double* src) {
#if defined(CB2_EMUL)
  *dst = *src;
#else
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(src));
#endif
}
CB2_D void cp_async_commit() {
#if !defined(CB2_EMUL)
  asm volatile("cp.async.commit_group;\n" ::);
#endif
}
template <int N>
CB2_D void cp_async_wait() {
#if !defined(CB2_EMUL)
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
#endif
}

// NB = number of 8-column blocks of the local layout, CAL0 = first calibration column: (6, 37) covers sensors with up to 11 calibration
// unknowns — the IMU models with 4 intrinsics: 21 DMMA tiles per k-step instead of 28 —, (7, 40) up to 16, (8, 40) up to 20 (kMaxCalib).
// Sensors whose Gram slots come from the sweep (sd.gslots != nullptr) are skipped here; the warps are dealt the remaining sensors.
template <int NB, int CAL0>
__global__ void __launch_bounds__(kAccThreads, (NB <= 7 ? CB2_ACC_MINBLOCKS : 2)) accumulate_kernel(
    const SensorDesc* __restrict__ sensors, const int* __restrict__ plain_idx, int n_plain, int n_local_seg, int spc, int N_c, int g_lo,
    const int* __restrict__ c2off, int csz, double* __restrict__ segA, double* __restrict__ segG, double* __restrict__ segB,
    double* __restrict__ segC, double* __restrict__ segGc) {
  // dynamic shared memory: per warp two tile buffers of kAccRows x kAccStride (+ 4 doubles: the last fragment reads past a row end)
  typedef double TileBuf[2][kAccRows * kAccStride + 4];
  TileBuf* tiles = dyn_smem<TileBuf>();
  __shared__ int colpos[kAccWarps][kCpCols + kMaxCalib];
  // spc segments per CTA (1, 2 or 4), kAccWarps / spc warps per segment: with fewer sensors than warps (C4: gyroscope + accelerometer) every
  // warp still has a (segment, sensor) pair of its own instead of half the CTA waiting at the final barrier.
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int wps = kAccWarps / spc, lw = warp % wps;
  const int gl = blockIdx.x * spc + warp / wps, g = g_lo + gl;
  const bool live = gl < n_local_seg;
  const int fr = lane & 3, fc = lane >> 2;   // fragment row (k) and column of this lane
  double acc[NB][NB][2];                     // [bi][bj <= bi]: 8x8 tile of the local Gram matrix; the rest is dead code
#pragma unroll
  for (int bi = 0; bi < NB; ++bi)
#pragma unroll
    for (int bj = 0; bj < NB; ++bj) { acc[bi][bj][0] = 0.0; acc[bi][bj][1] = 0.0; }
  double* const tb0 = tiles[warp][0];
  for (int i = lane; i < 2 * (kAccRows * kAccStride + 4); i += 32) tb0[i] = 0.0;   // padding columns stay zero for the whole kernel
  __syncwarp();
  // The sensors this kernel handles (plain_idx: those without Gram slots) are dealt round-robin to the warps of the segment.
  for (int k = lw; live && k < n_plain; k += wps) {
    const int s = plain_idx[k];
    const SensorDesc& sd = sensors[s];
    const int m = sd.m, jw = sd.jw, nc = sd.n_calib;
    const int o0 = sd.seg_start[g];
    const int rows = (sd.seg_start[g + 1] - o0) * m;
    if (rows == 0) continue;
    // Local column of every stored J column. The previous sensor's calibration columns are cleared first.
    if (lane < kAccStride - CAL0)
#pragma unroll 4
      for (int rr = 0; rr < 2 * kAccRows; ++rr) tb0[(rr / kAccRows) * (kAccRows * kAccStride + 4) + (rr % kAccRows) * kAccStride + CAL0 + lane] = 0.0;
    for (int c = lane; c < jw; c += 32) colpos[warp][c] = c < kCpCols ? c : CAL0 + sd.junk[c - kCpCols];
    __syncwarp();
    const double* __restrict__ Jg = sd.J + size_t(o0) * m * jw;
    const double* __restrict__ rg = sd.r + size_t(o0) * m;
    const int ntiles = (rows + kAccRows - 1) / kAccRows;
    // Lane l always copies stored columns l and l + 32 (jw <= 56): their local positions are loop-invariant registers, a warp
    // instruction reads 256 contiguous bytes of one J row.
    const int cp0 = colpos[warp][min(lane, jw - 1)], cp1 = colpos[warp][min(lane + 32, jw - 1)];
    const bool has1 = lane + 32 < jw;
    auto issue_load = [&](int tile, int buf) {
      if (tile < ntiles) {
        double* tb = tiles[warp][buf];
        const int r0 = tile * kAccRows;
        const int nr = min(kAccRows, rows - r0);
        const double* src = Jg + size_t(r0) * jw + lane;
#pragma unroll 4
        for (int rr = 0; rr < nr; ++rr) {
          cp_async8(tb + rr * kAccStride + cp0, src + rr * jw);
          if (has1) cp_async8(tb + rr * kAccStride + cp1, src + rr * jw + 32);
        }
        if (lane < kAccRows) {
          if (lane < nr) cp_async8(tb + lane * kAccStride + kAccRcol, rg + r0 + lane);
          else for (int c = 0; c < kAccStride; ++c) tb[lane * kAccStride + c] = 0.0;     // rows past the end of the segment
        }
      }
      cp_async_commit();
    };
    issue_load(0, 0);
    for (int tile = 0; tile < ntiles; ++tile) {
      const int buf = tile & 1;
      issue_load(tile + 1, buf ^ 1);
      cp_async_wait<1>();
      __syncwarp();
      const double* tb = tiles[warp][buf];
      const int nr = min(kAccRows, rows - tile * kAccRows);
#pragma unroll
      for (int ks = 0; ks < kAccRows / 4; ++ks) {
        if (4 * ks < nr) {
          double f[NB];
#pragma unroll
          for (int b = 0; b < NB; ++b) f[b] = tb[(4 * ks + fr) * kAccStride + 8 * b + fc];
#pragma unroll
          for (int bi = 0; bi < NB; ++bi)
#pragma unroll
            for (int bj = 0; bj <= bi; ++bj) dmma_8x8x4(acc[bi][bj][0], acc[bi][bj][1], f[bi], f[bj]);
        }
      }
      __syncwarp();
    }
    cp_async_wait<0>();
    // Flush every entry whose ROW is one of this sensor's calibration unknowns (local rows CAL0..), then reset them for the next sensor.
#pragma unroll
    for (int bi = CAL0 / 8; bi < NB; ++bi) {
      const int li = 8 * bi + fc - CAL0;      // calibration-local unknown of this lane's accumulator row; < 0: a control-point / residual row
#pragma unroll
      for (int bj = 0; bj <= bi; ++bj) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int Jx = 8 * bj + 2 * fr + i;
          const double v = acc[bi][bj][i];
          if (li >= 0 && li < nc) {
            if (Jx < kCpCols) segB[(size_t(gl) * kCpCols + Jx) * N_c + sd.calib_off + li] = v;
            else if (Jx == kAccRcol) segGc[size_t(gl) * N_c + sd.calib_off + li] = v;
            else if (Jx >= CAL0 && Jx - CAL0 <= li) segC[size_t(gl) * csz + c2off[s] + li * nc + (Jx - CAL0)] = v;
          }
          if (li >= 0) acc[bi][bj][i] = 0.0;
        }
      }
    }
  }
  // Control-point block (+ gradient row 36): combine the warps in a fixed order through shared memory.
  __syncwarp();
  double* red = tiles[warp][0];             // 15 tiles x 64 doubles = 960 <= 2 * (16 * 68 + 4)
  {
    int tix = 0;
#pragma unroll
    for (int bi = 0; bi < 5; ++bi)
#pragma unroll
      for (int bj = 0; bj <= bi; ++bj) {
        red[tix * 64 + fc * 8 + 2 * fr] = acc[bi][bj][0];
        red[tix * 64 + fc * 8 + 2 * fr + 1] = acc[bi][bj][1];
        ++tix;
      }
  }
  __syncthreads();
  // tix -> (bi, bj) of the lower block triangle without a search loop: tix = bi (bi + 1) / 2 + bj, bi < 5.
  for (int e2 = t; e2 < spc * 15 * 64; e2 += kAccThreads) {
    const int q = e2 / (15 * 64), e = e2 - q * (15 * 64);
    const int gq = blockIdx.x * spc + q;                 // local segment of warps q wps .. (q + 1) wps - 1
    if (gq >= n_local_seg) continue;
    const int tix = e >> 6, mrow = (e >> 3) & 7, ncol = e & 7;
    const int bi = tix >= 10 ? 4 : (tix >= 6 ? 3 : (tix >= 3 ? 2 : (tix >= 1 ? 1 : 0)));
    const int rem = tix - bi * (bi + 1) / 2;
    const int I = 8 * bi + mrow, Jx = 8 * rem + ncol;
    if (Jx > I || Jx >= kCpCols || I > kAccRcol) continue;
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kAccWarps; ++w) if (w < wps) v += tiles[q * wps + w][0][e];
    if (I < kCpCols) segA[(size_t(gq) * kCpCols + I) * kCpCols + Jx] = v;
    else segG[size_t(gq) * kCpCols + Jx] = v;
  }
}

// The camera part of the normal equations from the compact Gram slots the sweep left (see the header comment and SensorDesc::gslots).
// One CTA per spline segment, one WARP per camera (round-robin): a warp walks its camera's slots of this segment on its own — no block
// barrier until the end, so the dependent global loads (slot range -> image index -> basis weights -> slot data) of the warps overlap.
// Slots are staged through per-warp shared memory with coalesced 16-byte loads; the expansion then follows the Kronecker structure
//   H[cp_a, cp_b] += w_a w_b G (6 x 6):   lane l owns entry l of G for each of the 21 sub-blocks (a >= b) -> one load of G per slot, the six
//                                         weights broadcast, 21 FMAs; the 4 entries 32..35 of every sub-block and the gradient row
//                                         (w_b Gr) are "extra" entries, 4 per lane, described by ext_tab = (a << 16) | (b << 8) | offset;
//   H[cp_a, calib]  += w_a Gc (16 x 6):   lane l owns entries l, l + 32, l + 64 of Gc for each a -> 18 accumulators, written per camera.
// No atomics, fixed summation order. The warps' control-point partials are combined through shared memory at the end.
// The calibration x calibration blocks and the calibration gradient need no basis weights: the sweep sums them per (CTA, warp)
// (SensorDesc::gcta) and assemble_calib_kernel reduces them. The control-point block goes to its own per-segment buffer (accumulate_kernel
// runs beside this kernel on a second stream); assemble_band_kernel adds the two.
constexpr int kExpWarps = 8;
constexpr int kExpThreads = 32 * kExpWarps;
constexpr int kExpBatch = 4;                         // slots a warp stages at a time
constexpr int kExpSub = 21;                          // 6 x 6 sub-blocks (a >= b) of the control-point block
constexpr int kExpPart = kExpSub * 36 + 36;          // per-warp partial: 21 sub-blocks + the gradient row [b][q]
constexpr int kExpExt = 4;                           // extra entries per lane: 21 x 4 leftovers + 36 gradient entries = 120 <= 128
constexpr int kExpPerThread = (kExpPart + kExpThreads - 1) / kExpThreads;
static_assert(kExpWarps * kExpBatch * kGramSlot >= (kExpWarps / 2) * kExpPart, "the staging area doubles as the reduction buffer");
CB2_HD constexpr size_t expand_smem_bytes() { return size_t(kExpWarps) * kExpBatch * kGramSlot * sizeof(double); }
// ext_tab[lane * kExpExt + j] = (a << 16) | (b << 8) | offset of the slot entry, a == 6: gradient row (weight 1); -1 = unused.
// ext_dst[...] = index in the per-warp partial.
inline void expand_ext_table(int* tab, int* dst) {
  int n = 0;
  for (int i = 0; i < 32 * kExpExt; ++i) { tab[i] = -1; dst[i] = 0; }
  for (int a = 0; a < 6; ++a) for (int b = 0; b <= a; ++b) for (int e = 32; e < 36; ++e) { tab[n] = (a << 16) | (b << 8) | e; dst[n] = (a * (a + 1) / 2 + b) * 36 + e; ++n; }
  for (int b = 0; b < 6; ++b) for (int q = 0; q < 6; ++q) { tab[n] = (6 << 16) | (b << 8) | (36 + q); dst[n] = kExpSub * 36 + b * 6 + q; ++n; }
}
// slot_tab[(segment * n_gram + j)] = {first slot, slot count} of the j-th slot-producing camera in that segment and gram_meta[j] =
// {calib_off, n_calib} (both built by the host at upload): a warp reaches its slots through ONE table load instead of a chain of
// descriptor -> CSR -> slot-index loads.
__global__ void __launch_bounds__(kExpThreads, 2) expand_gram_kernel(const int2* __restrict__ slot_tab, const int2* __restrict__ gram_meta, int n_gram,
                                                                     const double* __restrict__ gslots, int N_c,
                                                                     const int* __restrict__ ext_tab, const int* __restrict__ ext_dst,
                                                                     double* __restrict__ segA, double* __restrict__ segG, double* __restrict__ segB) {
  double* smem = dyn_smem<double>();
  const int gl = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  double* const s_slot = smem + size_t(warp) * kExpBatch * kGramSlot;                               // [kExpBatch][kGramSlot]
  int et[kExpExt];
#pragma unroll
  for (int j = 0; j < kExpExt; ++j) et[j] = ext_tab[lane * kExpExt + j];
  double va[kExpSub], ve[kExpExt];
#pragma unroll
  for (int i = 0; i < kExpSub; ++i) va[i] = 0.0;
#pragma unroll
  for (int j = 0; j < kExpExt; ++j) ve[j] = 0.0;
  for (int j = warp; j < n_gram; j += kExpWarps) {
    {
    const int2 rng = slot_tab[size_t(gl) * n_gram + j];
    const int lo = rng.x, n = rng.y;
    if (n == 0) continue;
    const int2 meta = gram_meta[j];
    const int nc = meta.y, calib_off = meta.x;
    const double* __restrict__ slots = gslots + size_t(lo) * kGramSlot;
    for (int k0 = 0; k0 < n; k0 += kExpBatch) {
      const int kn = min(kExpBatch, n - k0);
      __syncwarp();
      {
        // the slots of a camera are contiguous: one coalesced copy, all loads of a lane issued before its first store
        constexpr int NL = (kExpBatch * (kGramSlot / 2) + 31) / 32;
        double2 buf[NL];
        const double2* __restrict__ src = reinterpret_cast<const double2*>(slots + size_t(k0) * kGramSlot);
        const int tot = kn * (kGramSlot / 2);
#pragma unroll
        for (int i = 0; i < NL; ++i) buf[i] = src[min(lane + 32 * i, tot - 1)];
#pragma unroll
        for (int i = 0; i < NL; ++i) if (lane + 32 * i < tot) reinterpret_cast<double2*>(s_slot)[lane + 32 * i] = buf[i];
      }
      __syncwarp();
      // control points x control points + gradient
      for (int k = 0; k < kn; ++k) {
        const double* __restrict__ S = s_slot + k * kGramSlot;
        if (S[kGramSlotFlag] == 0.0) continue;                  // unused slot (stale data)
        const double* __restrict__ w = S + kGramSlotW;          // w[6] is the flag = 1.0: the weight of the gradient row
        double wv[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) wv[a] = w[a];
        const double g0 = S[lane];                              // entry `lane` of [g | r] x g: row lane / 6, column lane % 6
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int b = 0; b <= a; ++b) va[a * (a + 1) / 2 + b] += (wv[a] * wv[b]) * g0;
#pragma unroll
        for (int j = 0; j < kExpExt; ++j)
          if (et[j] >= 0) ve[j] += w[et[j] >> 16] * w[(et[j] >> 8) & 255] * S[et[j] & 255];
      }
      // control points x calibration: entry e = lane + 32 j of Gc [16][6] (li = e / 6, p = e % 6), j < 3, for each a < 6
      double vb[18];
#pragma unroll
      for (int j = 0; j < 18; ++j) vb[j] = 0.0;
      for (int k = 0; k < kn; ++k) {
        const double* __restrict__ S = s_slot + k * kGramSlot;
        if (S[kGramSlotFlag] == 0.0) continue;
        const double c0 = S[48 + lane], c1 = S[48 + 32 + lane], c2 = S[48 + 64 + lane];
#pragma unroll
        for (int a = 0; a < 6; ++a) { const double wa = S[kGramSlotW + a]; vb[3 * a] += wa * c0; vb[3 * a + 1] += wa * c1; vb[3 * a + 2] += wa * c2; }
      }
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int e = lane + 32 * j, li = e / 6, p6 = e - 6 * li;
          if (li < nc) {
            double* dst = segB + (size_t(gl) * kCpCols + 6 * a + p6) * N_c + calib_off + li;
            *dst = k0 == 0 ? vb[3 * a + j] : *dst + vb[3 * a + j];   // more than kExpBatch slots of one camera in one segment: rare
          }
        }
    }
    }
  }
  // fixed-order combination of the warps' partials in two rounds through the (now free) staging area: warp w + 4 -> warp w, then the four sums
  double* const s_part = smem;                                      // [kExpWarps / 2][kExpPart]
  auto put = [&](double* dst, bool add) {
#pragma unroll
    for (int i = 0; i < kExpSub; ++i) dst[i * 36 + lane] = add ? dst[i * 36 + lane] + va[i] : va[i];
#pragma unroll
    for (int j = 0; j < kExpExt; ++j) if (et[j] >= 0) { const int d = ext_dst[lane * kExpExt + j]; dst[d] = add ? dst[d] + ve[j] : ve[j]; }
  };
  __syncthreads();
  if (warp >= kExpWarps / 2) put(s_part + size_t(warp - kExpWarps / 2) * kExpPart, false);
  __syncthreads();
  if (warp < kExpWarps / 2) put(s_part + size_t(warp) * kExpPart, true);
  __syncthreads();
  for (int k = t; k < kExpPart; k += kExpThreads) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kExpWarps / 2; ++w) v += s_part[size_t(w) * kExpPart + k];
    if (k < kExpSub * 36) {
      const int sub = k / 36, e = k - 36 * sub;
      int a = 0;
      while ((a + 1) * (a + 2) / 2 <= sub) ++a;
      const int bb = sub - a * (a + 1) / 2, p6 = e / 6, q6 = e - 6 * p6;
      const int I = 6 * a + p6, Jx = 6 * bb + q6;
      if (Jx <= I) segA[(size_t(gl) * kCpCols + I) * kCpCols + Jx] = v;   // (upper part of a diagonal sub-block: nothing to write)
    } else {
      segG[size_t(gl) * kCpCols + (k - kExpSub * 36)] = v;
    }
  }
}

// Banded A (lower band, A(i,j) at Aband[i*36 + 35 - (i-j)]), dense border Bmat[6 n_cp][N_c] and the control-point part of
// the gradient, from the per-segment partials of this rank's segments [g_lo, g_hi). Control point c belongs to segments c-5..c.
// segA2 / segG2: the cameras' control-point partials from expand_gram_kernel (nullptr when only one of the two kernels ran).
__global__ void __launch_bounds__(256) assemble_band_kernel(int n_cp, int g_lo, int g_hi, int N_c, const double* __restrict__ segA,
                                                            const double* __restrict__ segG, const double* __restrict__ segB,
                                                            const double* __restrict__ segA2, const double* __restrict__ segG2,
                                                            double* __restrict__ Aband, double* __restrict__ Bmat, double* __restrict__ grad) {
  const long n = 6L * n_cp;
  const long nA = n * kCpCols, nB = n * N_c;
  const long total = nA + nB + n;
  for (long idx = long(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += long(gridDim.x) * blockDim.x) {
    if (idx < nA) {
      const int i = int(idx / kCpCols), d = int(idx % kCpCols);
      const int j = i - (kCpCols - 1 - d);
      double s = 0.0;
      if (j >= 0) {
        const int ci = i / 6, cj = j / 6;
        const int g0 = max(ci - 5, g_lo), g1 = min(cj, g_hi - 1);
        for (int g = g0; g <= g1; ++g) s += segA[(size_t(g - g_lo) * kCpCols + (i - 6 * g)) * kCpCols + (j - 6 * g)];
        if (segA2) for (int g = g0; g <= g1; ++g) s += segA2[(size_t(g - g_lo) * kCpCols + (i - 6 * g)) * kCpCols + (j - 6 * g)];
      }
      Aband[idx] = s;
    } else if (idx < nA + nB) {
      const long e = idx - nA;
      const int i = int(e / N_c), c = int(e % N_c);
      const int ci = i / 6;
      const int g0 = max(ci - 5, g_lo), g1 = min(ci, g_hi - 1);
      double s = 0.0;
      for (int g = g0; g <= g1; ++g) s += segB[(size_t(g - g_lo) * kCpCols + (i - 6 * g)) * N_c + c];
      Bmat[e] = s;
    } else {
      const int i = int(idx - nA - nB);
      const int ci = i / 6;
      const int g0 = max(ci - 5, g_lo), g1 = min(ci, g_hi - 1);
      double s = 0.0;
      for (int g = g0; g <= g1; ++g) s += segG[size_t(g - g_lo) * kCpCols + (i - 6 * g)];
      if (segG2) for (int g = g0; g <= g1; ++g) s += segG2[size_t(g - g_lo) * kCpCols + (i - 6 * g)];
      grad[i] = s;
    }
  }
}

// Calibration block C (dense N_c x N_c storage, symmetric fill) and calibration gradient: reduction over all partials in two
// deterministic stages. An entry sums `count` partials `stride` doubles apart, starting at offset `src` of one of three buffers:
// kind 0 per-segment segC, kind 1 per-segment segGc (accumulate_kernel's sensors), kind 2 the per-CTA calibration tiles the sweep
// left for the cameras whose Gram is formed there (SensorDesc::gcta). Stage 1: grid = (entries / 32, kCalibSlices), blockDim = (32, 8):
// x = entry, (blockIdx.y, y) = slice of the partials; partial[slice][entry]. Stage 2: one thread per entry sums the slices in a fixed
// order and scatters.
struct CalibEntry { long src; int stride, count, kind; int dst_row, dst_col; };   // dst_col < 0: gradient entry dst_row
constexpr int kCalibSlices = 16;
__global__ void __launch_bounds__(256) assemble_calib_kernel(int n_entries, const CalibEntry* __restrict__ entries, const double* __restrict__ segC,
                                                             const double* __restrict__ segGc, const double* __restrict__ gcta,
                                                             double* __restrict__ partial) {
  __shared__ double part[8][33];
  const int e = blockIdx.x * 32 + threadIdx.x;
  const int g0 = blockIdx.y * 8 + threadIdx.y, gs = 8 * gridDim.y;
  double s = 0.0;
  if (e < n_entries) {
    const CalibEntry ce = entries[e];
    const double* __restrict__ base = (ce.kind == 0 ? segC : (ce.kind == 1 ? segGc : gcta)) + ce.src;
    int g = g0;
    for (; g + 3 * gs < ce.count; g += 4 * gs) {        // 4 independent loads in flight, summed in index order
      const double v0 = base[size_t(g) * ce.stride], v1 = base[size_t(g + gs) * ce.stride], v2 = base[size_t(g + 2 * gs) * ce.stride],
                   v3 = base[size_t(g + 3 * gs) * ce.stride];
      s += v0; s += v1; s += v2; s += v3;
    }
    for (; g < ce.count; g += gs) s += base[size_t(g) * ce.stride];
  }
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && e < n_entries) {
    double tot = 0.0;
    for (int y = 0; y < 8; ++y) tot += part[y][threadIdx.x];
    partial[size_t(blockIdx.y) * n_entries + e] = tot;
  }
}
__global__ void __launch_bounds__(256) assemble_calib_final_kernel(int N_c, int n_entries, int n_slices, const CalibEntry* __restrict__ entries,
                                                                   const double* __restrict__ partial, double* __restrict__ Cmat, double* __restrict__ grad_c) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_entries) return;
  const CalibEntry ce = entries[e];
  double tot = 0.0;
  for (int y = 0; y < n_slices; ++y) tot += partial[size_t(y) * n_entries + e];
  if (ce.dst_col >= 0) { Cmat[size_t(ce.dst_row) * N_c + ce.dst_col] = tot; Cmat[size_t(ce.dst_col) * N_c + ce.dst_row] = tot; }
  else grad_c[ce.dst_row] = tot;
}

}  // namespace cb2
