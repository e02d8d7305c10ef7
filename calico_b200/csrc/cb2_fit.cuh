// calico_b200 — trajectory spline FIT on the device: the step immediately before the hot path.
//
// Replaces BSpline<6, double>::FitSpline (reference calico/bspline.hpp:247-297): the reference forms the dense N_data x N_cp
// matrix X of basis weights, X^T X (N_cp x N_cp) and solves by column-pivoted Householder QR — O(N_cp^3), and its own TODO
// (bspline.hpp:287-289) notes that X^T X is banded, symmetric, positive definite. Here the same normal equations
//     (X^T X) C = X^T D,     X[j, seg(j) .. seg(j)+5] = U(t_j) M_seg(j)      (bspline.hpp:252-279)
// are built per spline segment and solved by a banded Cholesky of half-bandwidth k-1 = 5 with the 6 right-hand sides
// (the 6 pose dimensions share X) riding along.
//   fit_accumulate_kernel   one warp per segment: w = U M for every sample of the segment, local Gram w w^T (21 entries) and
//                           w (x) d (36 entries), fixed-order shuffle reduction -> per-segment partials (no atomics)
//   fit_assemble_kernel     band of X^T X (n_cp x 6) and X^T D (n_cp x 6) from the <= 6 overlapping segments per control point
//   fit_solve_kernel        one warp: row-oriented banded Cholesky with a 6-row ring in shared memory, forward substitution in the same
//                           sweep, backward substitution. Lane c < 6 owns right-hand side c. Inherently sequential in the control points.
#pragma once
#include "cb2_device.cuh"

namespace cb2 {

constexpr int kFitGram = 21;   // lower triangle of the 6x6 local Gram matrix, (a, b <= a) at a (a + 1) / 2 + b

__global__ void __launch_bounds__(128) fit_accumulate_kernel(int n_seg, const int* __restrict__ seg_start, const double* __restrict__ times,
                                                             const double* __restrict__ data, const double* __restrict__ knots,
                                                             const double* __restrict__ basis, double* __restrict__ segGram,
                                                             double* __restrict__ segRhs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x * 4 + warp;
  if (g >= n_seg) return;
  const double knot0 = knots[g + kK - 1], knot1 = knots[g + kK];
  const double* __restrict__ M = basis + size_t(g) * (kK * kK);
  double gram[kFitGram], rhs[kK * 6];
#pragma unroll
  for (int i = 0; i < kFitGram; ++i) gram[i] = 0.0;
#pragma unroll
  for (int i = 0; i < kK * 6; ++i) rhs[i] = 0.0;
  for (int j = seg_start[g] + lane; j < seg_start[g + 1]; j += 32) {
    double w[1][kK];
    spline_weights<1>(M, knot0, knot1, times[j], w);
    double d[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) d[c] = data[size_t(j) * 6 + c];
#pragma unroll
    for (int a = 0; a < kK; ++a) {
#pragma unroll
      for (int b = 0; b <= a; ++b) gram[a * (a + 1) / 2 + b] += w[0][a] * w[0][b];
#pragma unroll
      for (int c = 0; c < 6; ++c) rhs[a * 6 + c] += w[0][a] * d[c];
    }
  }
#pragma unroll
  for (int i = 0; i < kFitGram; ++i) {
    double v = gram[i];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) segGram[size_t(g) * kFitGram + i] = v;
  }
#pragma unroll
  for (int i = 0; i < kK * 6; ++i) {
    double v = rhs[i];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) segRhs[size_t(g) * (kK * 6) + i] = v;
  }
}

// Band of X^T X and the right-hand sides from the per-segment partials: control point i receives contributions from the segments
// g = i-5 .. i (local index a = i - g). Aband[i][d] = (X^T X)(i, i - d), Bfit[i][c] = (X^T D)(i, c).
__global__ void __launch_bounds__(256) fit_assemble_kernel(int n_cp, int n_seg, const double* __restrict__ segGram, const double* __restrict__ segRhs,
                                                           double* __restrict__ Aband, double* __restrict__ Bfit) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_cp * 12; e += gridDim.x * blockDim.x) {
    const int i = e / 12, q = e % 12;
    double s = 0.0;
    if (q < 6) {
      const int j = i - q;
      if (j >= 0) {
        const int g0 = max(i - (kK - 1), 0), g1 = min(j, n_seg - 1);
        for (int g = g0; g <= g1; ++g) { const int a = i - g, b = j - g; s += segGram[size_t(g) * kFitGram + a * (a + 1) / 2 + b]; }
      }
      Aband[size_t(i) * 6 + q] = s;
    } else {
      const int c = q - 6;
      const int g0 = max(i - (kK - 1), 0), g1 = min(i, n_seg - 1);
      for (int g = g0; g <= g1; ++g) s += segRhs[size_t(g) * (kK * 6) + (i - g) * 6 + c];
      Bfit[size_t(i) * 6 + c] = s;
    }
  }
}

// One warp. Lband[i][d] = L(i, i - d), d = 0..5; ctrl holds the forward-substituted right-hand sides, then the control points.
// Lane l < 6 fetches band entry d = l and right-hand side c = l of the NEXT row while the current one is factored.
__global__ void __launch_bounds__(32) fit_solve_kernel(int n_cp, const double* __restrict__ Aband, const double* __restrict__ Bfit, double rel_eps,
                                                       double* __restrict__ Lband, double* __restrict__ ctrl, int* __restrict__ fail) {
  __shared__ double ring[kK][kK + 1];   // ring[i % 6][d] = L(i, i - d) of the last 6 rows; [kK] = 1 / L(i, i)
  const int lane = threadIdx.x, l6 = min(lane, 5);
  // Regularisation scale: control points without supporting samples make X^T X singular (the reference's pivoted QR then returns
  // a basic solution); a relative 1e-14 on the diagonal keeps the factorisation defined and does not move supported control points.
  double dmax = 0.0;
  for (int i = lane; i < n_cp; i += 32) dmax = fmax(dmax, Aband[size_t(i) * 6]);
  for (int off = 16; off > 0; off >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, off));
  const double eps = fmax(1.0, dmax) * rel_eps;
  int bad = 0;
  double a_next = n_cp > 0 ? Aband[l6] : 0.0, b_next = n_cp > 0 ? Bfit[l6] : 0.0;
  double yh[kK - 1] = {0.0, 0.0, 0.0, 0.0, 0.0};   // forward-substituted values of rows i-1 .. i-5 of this lane's right-hand side
  for (int i = 0; i < n_cp; ++i) {
    const double a_cur = a_next, b_cur = b_next;
    if (i + 1 < n_cp) { a_next = Aband[size_t(i + 1) * 6 + l6]; b_next = Bfit[size_t(i + 1) * 6 + l6]; }
    // Row i of L (every lane computes the same values; lane 0 publishes them) and this lane's right-hand side.
    double Li[kK], Linv = 1.0;
#pragma unroll
    for (int d = kK - 1; d >= 0; --d) {
      const int j = i - d;
      double s = __shfl_sync(0xffffffffu, a_cur, d) + (d == 0 ? eps : 0.0);
      if (j >= 0) {
#pragma unroll
        for (int m = 1; m <= kK - 1; ++m) {          // common predecessors j - m: L(i, j-m) = Li[d + m], L(j, j-m) = row j of the ring
          if (d + m <= kK - 1 && j - m >= 0) s -= Li[d + m] * (d == 0 ? Li[m] : ring[j % kK][m]);
        }
        if (d > 0) s *= ring[j % kK][kK];
        else { if (!(s > 0.0) || !isfinite(s)) { bad = 1; s = 1.0; } Linv = rsqrt(s); s *= Linv; }
      } else s = 0.0;
      Li[d] = s;
    }
    {
      double y = b_cur;
#pragma unroll
      for (int d = 1; d <= kK - 1; ++d) y -= Li[d] * yh[d - 1];   // Li[d] = 0 for rows before the first
      y *= Linv;
#pragma unroll
      for (int d = kK - 2; d > 0; --d) yh[d] = yh[d - 1];
      yh[0] = y;
      if (lane < 6) ctrl[size_t(i) * 6 + lane] = y;
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int d = 0; d < kK; ++d) { ring[i % kK][d] = Li[d]; Lband[size_t(i) * kK + d] = Li[d]; }
      ring[i % kK][kK] = Linv;
    }
    __syncwarp();
  }
  // L^T C = Y, backwards.
  if (lane < 6) {
    for (int i = n_cp - 1; i >= 0; --i) {
      double s = ctrl[size_t(i) * 6 + lane];
#pragma unroll
      for (int d = 1; d <= kK - 1; ++d) if (i + d < n_cp) s -= Lband[size_t(i + d) * kK + d] * ctrl[size_t(i + d) * 6 + lane];
      ctrl[size_t(i) * 6 + lane] = s / Lband[size_t(i) * kK];
    }
  }
  if (lane == 0) *fail = bad;
}

}  // namespace cb2
