// calico_b200 — K4: block-sparse Gauss-Newton normal equations  H = J^T J,  g = J^T r.
//
// Replaces Ceres's BlockSparseMatrix / SchurEliminator accumulation (Ceres external; selected by
// batch_optimizer.cpp:12 DENSE_SCHUR). Unknown vector = [control points 6*n_cp | calibration N_c]. Structure
// (SURVEY §8a): A = H[cp,cp] is block-banded (6x6 blocks, half-bandwidth k-1 = 5 blocks, i.e. 35 scalars) because a
// residual touches k consecutive control points (camera_cost_functor.cpp:52-60); B = H[cp,calib] couples a segment to
// the sensors observed in it; C = H[calib,calib] is block-diagonal per sensor (no residual involves two sensors).
//
// accumulate_kernel: one CTA per spline segment, one WARP per sensor (round-robin). All residual rows of a segment touch the
// same 36 control-point columns, so each warp forms the local Gram matrix X^T X of X = [J_cp | r | 0 0 0 | J_calib | 0..]
// (r as an extra column gives the gradient for free) on the FP64 TENSOR pipe: mma.sync.m8n8k4.f64 (SASS DMMA), 8x8
// accumulator tiles in registers, lower block triangle only. For a Gram product the A and B fragments of a column block are
// the same register (lane holds X[4 ks + lane % 4][8 b + lane / 4]), so a k-step of 4 rows costs NB shared loads for
// NB (NB + 1) / 2 DMMAs. J row tiles stream global -> shared with cp.async (per-warp double buffer, row stride 68 doubles
// == 4 mod 16: conflict-free fragment loads). The control-point block accumulates over all sensors of the warp and is
// combined across warps in a fixed order at the end; the calibration blocks are flushed per sensor. No atomics.
// assemble_*_kernel: sums the <= 6 overlapping segment partials per control-point entry into the banded storage and
// reduces the calibration blocks over all segments.
#pragma once
#include "cb2_device.cuh"

namespace cb2 {

constexpr int kAccWarps = 4;
constexpr int kAccThreads = 32 * kAccWarps;
constexpr int kAccRows = 16;       // J rows per shared-memory tile (4 DMMA k-steps)
constexpr int kAccStride = 68;     // doubles per tile row; 68 mod 16 == 4 -> the 16 lanes of a half warp hit 16 distinct 8-byte banks
constexpr int kAccRcol = 36;       // local layout: cp 0..35 | r 36 | zeros 37..39 | calibration unknowns 40.. | zeros
constexpr int kAccCal0 = 40;
constexpr size_t kAccSmemBytes = size_t(kAccWarps) * 2 * (kAccRows * kAccStride + 4) * sizeof(double);
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

// D(8x8) += A(8x4) B(4x8) in FP64 on the tensor pipe. Lane l holds a = A[l / 4][l % 4], b = B[l % 4][l / 4],
// c0, c1 = D[l / 4][2 (l % 4) + {0, 1}].
CB2_D void dmma_8x8x4(double& c0, double& c1, double a, double b) {
#if defined(CB2_EMUL)
  ::cb2emul::dmma_8x8x4(c0, c1, a, b);
#else
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
#endif
}

// NB = number of 8-column blocks of the local layout: 7 covers sensors with up to 16 calibration unknowns, 8 up to 20 (kMaxCalib).
template <int NB>
__global__ void __launch_bounds__(kAccThreads, (NB == 7 ? CB2_ACC_MINBLOCKS : 2)) accumulate_kernel(
    const SensorDesc* __restrict__ sensors, int n_sensors, int N_c, int g_lo, const int* __restrict__ c2off, int csz, double* __restrict__ segA,
    double* __restrict__ segG, double* __restrict__ segB, double* __restrict__ segC, double* __restrict__ segGc) {
  // dynamic shared memory: per warp two tile buffers of kAccRows x kAccStride (+ 4 doubles: the last fragment reads past a row end)
  typedef double TileBuf[2][kAccRows * kAccStride + 4];
  TileBuf* tiles = dyn_smem<TileBuf>();
  __shared__ int colpos[kAccWarps][kCpCols + kMaxCalib];
  const int gl = blockIdx.x, g = g_lo + gl, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int fr = lane & 3, fc = lane >> 2;   // fragment row (k) and column of this lane
  double acc[NB][NB][2];                     // [bi][bj <= bi]: 8x8 tile of the local Gram matrix; the rest is dead code
#pragma unroll
  for (int bi = 0; bi < NB; ++bi)
#pragma unroll
    for (int bj = 0; bj < NB; ++bj) { acc[bi][bj][0] = 0.0; acc[bi][bj][1] = 0.0; }
  double* const tb0 = tiles[warp][0];
  for (int i = lane; i < 2 * (kAccRows * kAccStride + 4); i += 32) tb0[i] = 0.0;   // padding columns stay zero for the whole kernel
  __syncwarp();

  for (int s = warp; s < n_sensors; s += kAccWarps) {
    const SensorDesc& sd = sensors[s];
    const int m = sd.m, jw = sd.jw, nc = sd.n_calib;
    const int o0 = sd.seg_start[g];
    const int rows = (sd.seg_start[g + 1] - o0) * m;
    if (rows == 0) continue;
    // Local column of every stored J column. The previous sensor's calibration columns are cleared first.
    if (lane < kAccStride - kAccCal0)
#pragma unroll 4
      for (int rr = 0; rr < 2 * kAccRows; ++rr) tb0[(rr / kAccRows) * (kAccRows * kAccStride + 4) + (rr % kAccRows) * kAccStride + kAccCal0 + lane] = 0.0;
    for (int c = lane; c < jw; c += 32) colpos[warp][c] = c < kCpCols ? c : kAccCal0 + sd.junk[c - kCpCols];
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
    // Flush every entry that involves this sensor's calibration unknowns (blocks 5..NB-1), then reset them for the next sensor.
#pragma unroll
    for (int bi = 5; bi < NB; ++bi) {
      const int li = 8 * (bi - 5) + fc;       // calibration-local unknown of this lane's accumulator row
#pragma unroll
      for (int bj = 0; bj <= bi; ++bj) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int Jx = 8 * bj + 2 * fr + i;
          const double v = acc[bi][bj][i];
          if (li < nc) {
            if (Jx < kCpCols) segB[(size_t(gl) * kCpCols + Jx) * N_c + sd.calib_off + li] = v;
            else if (Jx == kAccRcol) segGc[size_t(gl) * N_c + sd.calib_off + li] = v;
            else if (Jx >= kAccCal0 && Jx - kAccCal0 <= li) segC[size_t(gl) * csz + c2off[s] + li * nc + (Jx - kAccCal0)] = v;
          }
          acc[bi][bj][i] = 0.0;
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
  for (int e = t; e < 15 * 64; e += kAccThreads) {
    const int tix = e >> 6, mrow = (e >> 3) & 7, ncol = e & 7;
    const int bi = tix >= 10 ? 4 : (tix >= 6 ? 3 : (tix >= 3 ? 2 : (tix >= 1 ? 1 : 0)));
    const int rem = tix - bi * (bi + 1) / 2;
    const int I = 8 * bi + mrow, Jx = 8 * rem + ncol;
    if (Jx > I || Jx >= kCpCols || I > kAccRcol) continue;
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kAccWarps; ++w) v += tiles[w][0][e];
    if (I < kCpCols) segA[(size_t(gl) * kCpCols + I) * kCpCols + Jx] = v;
    else segG[size_t(gl) * kCpCols + Jx] = v;
  }
}

// Banded A (lower band, A(i,j) at Aband[i*36 + 35 - (i-j)]), dense border Bmat[6 n_cp][N_c] and the control-point part of
// the gradient, from the per-segment partials of this rank's segments [g_lo, g_hi). Control point c belongs to segments c-5..c.
__global__ void __launch_bounds__(256) assemble_band_kernel(int n_cp, int g_lo, int g_hi, int N_c, const double* __restrict__ segA,
                                                            const double* __restrict__ segG, const double* __restrict__ segB,
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
      grad[i] = s;
    }
  }
}

// Calibration block C (dense N_c x N_c storage, symmetric fill) and calibration gradient: reduction over all segments in two
// deterministic stages. Stage 1: grid = (entries / 32, kCalibSlices), blockDim = (32, 8): x = entry, (blockIdx.y, y) = segment slice;
// partial[slice][entry]. Stage 2: one thread per entry sums the slices in a fixed order and scatters.
struct CalibEntry { int src; int dst_row, dst_col; };   // src: offset in a segment's segC (or segGc when dst_col < 0)
constexpr int kCalibSlices = 16;
__global__ void __launch_bounds__(256) assemble_calib_kernel(int n_seg, int N_c, int csz, int n_entries, const CalibEntry* __restrict__ entries,
                                                             const double* __restrict__ segC, const double* __restrict__ segGc,
                                                             double* __restrict__ partial) {
  __shared__ double part[8][33];
  const int e = blockIdx.x * 32 + threadIdx.x;
  const int g0 = blockIdx.y * 8 + threadIdx.y, gs = 8 * gridDim.y;
  double s = 0.0;
  if (e < n_entries) {
    const CalibEntry ce = entries[e];
    if (ce.dst_col >= 0) { for (int g = g0; g < n_seg; g += gs) s += segC[size_t(g) * csz + ce.src]; }
    else { for (int g = g0; g < n_seg; g += gs) s += segGc[size_t(g) * N_c + ce.src]; }
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
