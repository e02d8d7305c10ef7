// calico_b200 — K4: block-sparse Gauss-Newton normal equations  H = J^T J,  g = J^T r.
//
// Replaces Ceres's BlockSparseMatrix / SchurEliminator accumulation (Ceres external; selected by
// batch_optimizer.cpp:12 DENSE_SCHUR). Unknown vector = [control points 6*n_cp | calibration N_c]. Structure
// (SURVEY §8a): A = H[cp,cp] is block-banded (6x6 blocks, half-bandwidth k-1 = 5 blocks, i.e. 35 scalars) because a
// residual touches k consecutive control points (camera_cost_functor.cpp:52-60); B = H[cp,calib] couples a segment to
// the sensors observed in it; C = H[calib,calib] is block-diagonal per sensor (no residual involves two sensors).
//
// accumulate_kernel: one CTA per spline segment. All residual rows of a segment touch the same 36 control-point
// columns, so the CTA accumulates one local (36 + n_calib + 1)^2 Gram matrix per sensor ([J | r]^T [J | r], r as an
// extra column gives the gradient for free) in registers — 6x6 micro-tiles, lower triangle only (55 threads) — from J row
// tiles streamed into shared memory with cp.async (double-buffered, the next tile is in flight while the current one is
// multiplied; 16-byte shared loads), and writes per-segment partials with plain stores (no atomics; fixed summation order).
// assemble_*_kernel: sums the <= 6 overlapping segment partials per control-point entry into the banded storage and
// reduces the calibration blocks over all segments.
#pragma once
#include "cb2_device.cuh"

namespace cb2 {

constexpr int kAccThreads = 64;
#ifndef CB2_ACC_ROWS
#define CB2_ACC_ROWS 24
#endif
#ifndef CB2_ACC_MINBLOCKS
#define CB2_ACC_MINBLOCKS 8
#endif
constexpr int kAccRows = CB2_ACC_ROWS;   // J rows per shared-memory tile
constexpr int kAccW = 60;      // local width: 36 cp | <= 20 calib | r at position 56 | 3 pad  (10 tiles of 6)
constexpr int kAccRcol = 56;
constexpr int kAccTiles = 55;  // lower-triangular 6x6 tiles of a 10 x 10 tile grid

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

struct AccCursor { int s, r0; };

__global__ void __launch_bounds__(kAccThreads, CB2_ACC_MINBLOCKS) accumulate_kernel(const SensorDesc* __restrict__ sensors, int n_sensors, int N_c, int g_lo,
                                                                 const int* __restrict__ c2off, int csz, double* __restrict__ segA,
                                                                 double* __restrict__ segG, double* __restrict__ segB,
                                                                 double* __restrict__ segC, double* __restrict__ segGc) {
  __shared__ __align__(16) double tile[2][kAccRows * kAccW];
  const int gl = blockIdx.x, g = g_lo + gl, t = threadIdx.x;   // gl: index into this rank's partial buffers, g: segment
  const bool has_tile = t < kAccTiles;
  int ti = 0, tj = 0;
  if (has_tile) { int rem = t; while (rem > ti) { rem -= ti + 1; ++ti; } tj = rem; }
  double acc[6][6];
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int b = 0; b < 6; ++b) acc[a][b] = 0.0;

  auto rows_of = [&](int s) { const SensorDesc& sd = sensors[s]; return (sd.seg_start[g + 1] - sd.seg_start[g]) * sd.m; };
  auto skip_empty = [&](AccCursor c) { while (c.s < n_sensors && rows_of(c.s) == 0) ++c.s; return c; };
  auto next_tile = [&](AccCursor c) {
    c.r0 += kAccRows;
    if (c.r0 >= rows_of(c.s)) { c.r0 = 0; ++c.s; c = skip_empty(c); }
    return c;
  };
  // Thread t < 60 fills local column position t of every tile row: the J column that maps there, the residual, or zero.
  auto issue_load = [&](AccCursor c, int buf) {
    if (c.s < n_sensors && t < kAccW) {
      const SensorDesc& sd = sensors[c.s];
      const int m = sd.m, jw = sd.jw;
      const int o0 = sd.seg_start[g];
      const int rows = (sd.seg_start[g + 1] - o0) * m;
      const int nr = min(kAccRows, rows - c.r0);
      int src = -1;                       // J column feeding this position
      if (t < kCpCols) src = t;
      else if (t < kAccRcol) { for (int j = 0; j < sd.n_jcal; ++j) if (kCpCols + sd.junk[j] == t) src = kCpCols + j; }
      double* dst = tile[buf] + t;
      if (src >= 0) {
        const double* base = sd.J + (size_t(o0) * m + c.r0) * jw + src;
        for (int row = 0; row < nr; ++row) cp_async8(dst + row * kAccW, base + size_t(row) * jw);
      } else if (t == kAccRcol) {
        const double* base = sd.r + size_t(o0) * m + c.r0;
        for (int row = 0; row < nr; ++row) cp_async8(dst + row * kAccW, base + row);
      } else {
        for (int row = 0; row < nr; ++row) dst[row * kAccW] = 0.0;
      }
    }
    cp_async_commit();
  };

  AccCursor cu = skip_empty(AccCursor{0, 0});
  int buf = 0;
  issue_load(cu, buf);
  for (int s = 0; s < n_sensors; ++s) {
    const SensorDesc& sd = sensors[s];
    const int nc = sd.n_calib;
    while (cu.s == s) {
      const int nr = min(kAccRows, rows_of(s) - cu.r0);
      const AccCursor nx = next_tile(cu);
      issue_load(nx, buf ^ 1);
      cp_async_wait<1>();
      __syncthreads();
      if (has_tile) {
        const double* tb = tile[buf];
        for (int row = 0; row < nr; ++row) {
          const double2* pa = reinterpret_cast<const double2*>(tb + row * kAccW + 6 * ti);
          const double2* pb = reinterpret_cast<const double2*>(tb + row * kAccW + 6 * tj);
          const double2 a0 = pa[0], a1 = pa[1], a2 = pa[2], b0 = pb[0], b1 = pb[1], b2 = pb[2];
          const double a6[6] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y};
          const double b6[6] = {b0.x, b0.y, b1.x, b1.y, b2.x, b2.y};
#pragma unroll
          for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b < 6; ++b) acc[a][b] += a6[a] * b6[b];
        }
      }
      __syncthreads();
      cu = nx;
      buf ^= 1;
    }
    // Flush every entry that involves this sensor's calibration columns, then reset it for the next sensor.
    if (has_tile && ti >= 6) {
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = 0; b < 6; ++b) {
          const int I = 6 * ti + a, Jx = 6 * tj + b;
          if (I == kAccRcol) {                    // residual row
            if (Jx < kCpCols) continue;          // control-point gradient: accumulated over all sensors
            const int lc = Jx - kCpCols;
            if (Jx < kAccRcol && lc < nc) segGc[size_t(gl) * N_c + sd.calib_off + lc] = acc[a][b];
          } else if (I < kAccRcol) {              // calibration row
            const int li = I - kCpCols;
            if (li < nc) {
              if (Jx < kCpCols) segB[(size_t(gl) * kCpCols + Jx) * N_c + sd.calib_off + li] = acc[a][b];
              else if (Jx <= I) segC[size_t(gl) * csz + c2off[s] + li * nc + (Jx - kCpCols)] = acc[a][b];
            }
          }
          acc[a][b] = 0.0;
        }
    }
  }
  cp_async_wait<0>();
  if (has_tile) {
    if (ti < 6) {
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = 0; b < 6; ++b) {
          const int I = 6 * ti + a, Jx = 6 * tj + b;
          if (Jx <= I) segA[(size_t(gl) * kCpCols + I) * kCpCols + Jx] = acc[a][b];
        }
    } else if (ti == 9 && tj < 6) {
#pragma unroll
      for (int b = 0; b < 6; ++b) segG[size_t(gl) * kCpCols + 6 * tj + b] = acc[kAccRcol - 54][b];
    }
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

// Calibration block C (dense N_c x N_c storage, symmetric fill) and calibration gradient: reduction over all segments.
// blockDim = (32, 8): x = entry, y = segment slice; the 8 slices are combined in a fixed order.
struct CalibEntry { int src; int dst_row, dst_col; };   // src: offset in a segment's segC (or segGc when dst_col < 0)
__global__ void __launch_bounds__(256) assemble_calib_kernel(int n_seg, int N_c, int csz, int n_entries, const CalibEntry* __restrict__ entries,
                                                             const double* __restrict__ segC, const double* __restrict__ segGc,
                                                             double* __restrict__ Cmat, double* __restrict__ grad_c) {
  __shared__ double part[8][33];
  const int e = blockIdx.x * 32 + threadIdx.x;
  double s = 0.0;
  CalibEntry ce; ce.src = 0; ce.dst_row = 0; ce.dst_col = 0;
  if (e < n_entries) {
    ce = entries[e];
    if (ce.dst_col >= 0) { for (int g = threadIdx.y; g < n_seg; g += 8) s += segC[size_t(g) * csz + ce.src]; }
    else { for (int g = threadIdx.y; g < n_seg; g += 8) s += segGc[size_t(g) * N_c + ce.src]; }
  }
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && e < n_entries) {
    double tot = 0.0;
    for (int y = 0; y < 8; ++y) tot += part[y][threadIdx.x];
    if (ce.dst_col >= 0) { Cmat[size_t(ce.dst_row) * N_c + ce.dst_col] = tot; Cmat[size_t(ce.dst_col) * N_c + ce.dst_row] = tot; }
    else grad_c[ce.dst_row] = tot;
  }
}

}  // namespace cb2
