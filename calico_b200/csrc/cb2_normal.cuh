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
// extra column gives the gradient for free) in registers — 4x4 micro-tiles, lower triangle only — from J row tiles staged in
// shared memory, and writes per-segment partials with plain stores (no atomics; summation order is fixed).
// assemble_*_kernel: sums the <= 6 overlapping segment partials per control-point entry into the banded storage and
// reduces the calibration blocks over all segments.
#pragma once
#include "cb2_device.cuh"

namespace cb2 {

constexpr int kAccThreads = 128;
constexpr int kAccRows = 32;   // J rows per shared-memory tile
constexpr int kAccW = 60;      // local width: 36 cp | <= 20 calib | r at column 56 | 3 pad
constexpr int kAccRcol = 56;
constexpr int kAccTiles = 120; // lower-triangular 4x4 tiles of a 15 x 15 tile grid

__global__ void __launch_bounds__(kAccThreads) accumulate_kernel(const SensorDesc* __restrict__ sensors, int n_sensors, int N_c,
                                                                 const int* __restrict__ c2off, int csz, double* __restrict__ segA,
                                                                 double* __restrict__ segG, double* __restrict__ segB,
                                                                 double* __restrict__ segC, double* __restrict__ segGc) {
  __shared__ __align__(16) double tile[kAccRows * kAccW];
  const int g = blockIdx.x, t = threadIdx.x;
  const bool has_tile = t < kAccTiles;
  int ti = 0, tj = 0;
  if (has_tile) { int rem = t; while (rem > ti) { rem -= ti + 1; ++ti; } tj = rem; }
  double acc[4][4];
  for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  for (int s = 0; s < n_sensors; ++s) {
    const SensorDesc& sd = sensors[s];
    const int m = sd.m, jw = sd.jw, nc = sd.n_calib;
    const int o0 = sd.seg_start[g], o1 = sd.seg_start[g + 1];
    const int rows = (o1 - o0) * m;
    const double* __restrict__ Jbase = sd.J + size_t(o0) * m * jw;
    const double* __restrict__ rbase = sd.r + size_t(o0) * m;
    for (int r0 = 0; r0 < rows; r0 += kAccRows) {
      const int nr = min(kAccRows, rows - r0);
      __syncthreads();
      for (int e = t; e < kAccRows * kAccW; e += kAccThreads) tile[e] = 0.0;
      __syncthreads();
      for (int e = t; e < nr * jw; e += kAccThreads) {
        const int row = e / jw, j = e - row * jw;
        const int pos = j < kCpCols ? j : kCpCols + sd.junk[j - kCpCols];
        tile[row * kAccW + pos] = Jbase[size_t(r0) * jw + e];
      }
      for (int row = t; row < nr; row += kAccThreads) tile[row * kAccW + kAccRcol] = rbase[r0 + row];
      __syncthreads();
      if (has_tile) {
        for (int row = 0; row < nr; ++row) {
          const double* tr = tile + row * kAccW;
          double a4[4], b4[4];
          for (int a = 0; a < 4; ++a) { a4[a] = tr[4 * ti + a]; b4[a] = tr[4 * tj + a]; }
          for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) acc[a][b] += a4[a] * b4[b];
        }
      }
    }
    // Flush every tile that involves this sensor's calibration columns, then reset it for the next sensor.
    if (has_tile && (tj >= 9 || (ti >= 9 && ti < 14))) {
      for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) {
        const int I = 4 * ti + a, Jx = 4 * tj + b;
        if (ti == 14) {                       // r row x calibration column -> calibration gradient
          if (tj >= 9 && tj < 14 && a == 0) { const int lc = Jx - kCpCols; if (lc < nc) segGc[size_t(g) * N_c + sd.calib_off + lc] = acc[a][b]; }
        } else if (tj < 9) {                  // calibration row x control-point column
          const int lc = I - kCpCols;
          if (lc < nc) segB[(size_t(g) * kCpCols + Jx) * N_c + sd.calib_off + lc] = acc[a][b];
        } else {                              // calibration x calibration (lower)
          const int li = I - kCpCols, lj = Jx - kCpCols;
          if (li < nc && lj <= li) segC[size_t(g) * csz + c2off[s] + li * nc + lj] = acc[a][b];
        }
        if (!(ti == 14 && tj < 9)) acc[a][b] = 0.0;
      }
    }
  }
  if (has_tile && tj < 9) {
    if (ti < 9) {
      for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) {
        const int I = 4 * ti + a, Jx = 4 * tj + b;
        if (Jx <= I) segA[(size_t(g) * kCpCols + I) * kCpCols + Jx] = acc[a][b];
      }
    } else if (ti == 14) {
      for (int b = 0; b < 4; ++b) segG[size_t(g) * kCpCols + 4 * tj + b] = acc[0][b];
    }
  }
}

// Banded A (lower band, A(i,j) at Aband[i*36 + 35 - (i-j)]), dense border Bmat[6 n_cp][N_c] and the control-point part of
// the gradient, from the per-segment partials. Control point c belongs to segments c-5..c.
__global__ void __launch_bounds__(256) assemble_band_kernel(int n_cp, int n_seg, int N_c, const double* __restrict__ segA,
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
        const int g0 = max(ci - 5, 0), g1 = min(cj, n_seg - 1);
        for (int g = g0; g <= g1; ++g) s += segA[(size_t(g) * kCpCols + (i - 6 * g)) * kCpCols + (j - 6 * g)];
      }
      Aband[idx] = s;
    } else if (idx < nA + nB) {
      const long e = idx - nA;
      const int i = int(e / N_c), c = int(e % N_c);
      const int ci = i / 6;
      const int g0 = max(ci - 5, 0), g1 = min(ci, n_seg - 1);
      double s = 0.0;
      for (int g = g0; g <= g1; ++g) s += segB[(size_t(g) * kCpCols + (i - 6 * g)) * N_c + c];
      Bmat[e] = s;
    } else {
      const int i = int(idx - nA - nB);
      const int ci = i / 6;
      const int g0 = max(ci - 5, 0), g1 = min(ci, n_seg - 1);
      double s = 0.0;
      for (int g = g0; g <= g1; ++g) s += segG[size_t(g) * kCpCols + (i - 6 * g)];
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
