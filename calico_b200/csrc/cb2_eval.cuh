// calico_b200 — K1/K2/K3 (residual + analytic Jacobian sweep), K8 (cost only) and K9 (final residuals).
//
// Replaces, for every residual block at once, what Ceres does per block through
// DynamicAutoDiffCostFunction::Evaluate -> {Camera,Gyroscope,Accelerometer}CostFunctor::operator()<Jet>
// (reference calico/sensors/camera_cost_functor.h:72-147, gyroscope_cost_functor.h:59-118,
// accelerometer_cost_functor.h:63-147), the loss corrector (Ceres external; for Huber/Cauchy rho'' <= 0 so residual and
// Jacobian are both scaled by sqrt(rho'), SURVEY §8 trap 6) and, in kModeResiduals, Sensor::UpdateResiduals
// (camera.cpp:70-80).
//
// One CTA = one tile of kTile consecutive observations of one sensor (observations are sorted by spline segment, so a
// tile touches one or two segments' control points). Phase 1: one thread per residual block evaluates the functor and
// leaves a compact derivative record in shared memory, field-major ([field][tile lane] — conflict free). Phase 2: one
// warp per residual block expands the record into the m x jw Jacobian rows and streams them to HBM with fully
// coalesced 8-byte stores (consecutive lanes -> consecutive doubles of one row).
#pragma once
#include "cb2_device.cuh"

namespace cb2 {

enum { kModeCost = 0, kModeResiduals = 1, kModeJacobian = 2 };

#ifndef CB2_EVAL_MINBLOCKS
#define CB2_EVAL_MINBLOCKS 8         // one-warp IMU CTAs: 8 per SM = up to 255 registers per thread
#endif
#ifndef CB2_EVAL_MINBLOCKS_CAM
#define CB2_EVAL_MINBLOCKS_CAM 4      // 4 x (52 fields x 129 x 8 B + static) = 224 KB of the SM's 228 KB
#endif
template <int KIND, int MODE>
__global__ void __launch_bounds__(eval_tile(KIND), (KIND == kCamera ? CB2_EVAL_MINBLOCKS_CAM : CB2_EVAL_MINBLOCKS)) eval_kernel(const SensorDesc* __restrict__ sensors, const SensorState* __restrict__ states,
                                                     const EvalTile* __restrict__ tiles, const double* __restrict__ ctrl,
                                                     const double* __restrict__ knots, const double* __restrict__ basis,
                                                     const double* __restrict__ pw, const double* __restrict__ frames, double gx, double gy, double gz,
                                                     double* __restrict__ cost_partial, int* __restrict__ invalid_partial, int apply_loss) {
  constexpr int kTile = eval_tile(KIND), kRecStride = eval_rec_stride(KIND);
  double* rec = dyn_smem<double>();
  __shared__ double s_red[kTile / 32];
  __shared__ int s_bad[kTile / 32];
  __shared__ unsigned char s_ok[kTile];
  constexpr int m = (KIND == kCamera) ? 2 : 3;
  const EvalTile tl = tiles[blockIdx.x];
  const SensorDesc& sd = sensors[tl.sensor];
  const int t = threadIdx.x;
  const bool active = t < tl.count;
  double cost = 0.0;
  int bad = 0;
  bool ok = false;
  double pc_z = 1.0;   // cameras, kModeResiduals: depth of the point in the camera frame (flag 2 of `valid`)
  if (active) {
    const SensorState S = states[tl.sensor];
    const long o = long(tl.start) + t;
    const Rec rc{rec + t, kRecStride};
    if (KIND == kCamera) {
      // Everything that depends only on (camera, stamp) was computed once per image by camera_frame_kernel.
      const int p = sd.pt[o];
      const double2 px = reinterpret_cast<const double2*>(sd.meas)[o];
      ok = camera_block_from_frame<MODE == kModeJacobian>(S, frames + size_t(sd.frame_base + sd.frm[o]) * FrameRec::kSize, px.x, px.y,
                                                          v3(pw[3 * p], pw[3 * p + 1], pw[3 * p + 2]), rc, MODE == kModeResiduals ? &pc_z : nullptr);
    } else {
      const double stamp = sd.stamp[o];
      const int seg = sd.seg[o];
      const double knot0 = knots[seg + kK - 1], knot1 = knots[seg + kK];
      const double* M = basis + size_t(seg) * (kK * kK);
      const double* cp = ctrl + size_t(seg) * 6;
      if (KIND == kGyroscope) {
        ok = gyro_block<MODE == kModeJacobian>(S, M, knot0, knot1, cp, stamp, v3(sd.meas[3 * o], sd.meas[3 * o + 1], sd.meas[3 * o + 2]), rc);
      } else {
        ok = accel_block<MODE == kModeJacobian>(S, v3(gx, gy, gz), M, knot0, knot1, cp, stamp,
                                                v3(sd.meas[3 * o], sd.meas[3 * o + 1], sd.meas[3 * o + 2]), rc);
      }
    }
    if (ok) {
      double sq = 0.0;
      for (int q = 0; q < m; ++q) { const double v = rc.get(q); sq += v * v; }
      double rho0, rho1;
      loss_eval(S.loss_type, S.loss_scale, sq, &rho0, &rho1);
      cost = 0.5 * rho0;
      if (MODE == kModeJacobian) {
        // Loss corrector (rho'' <= 0 for Huber/Cauchy): residual and Jacobian rows are both scaled by sqrt(rho'). Folded
        // into the derivative fields here so that the expansion phase is a pure sum of products.
        const double rs = apply_loss ? sqrt(rho1) : 1.0;
        if (rs != 1.0) {
          constexpr int d0 = (KIND == kCamera) ? int(CamRec::G0) : (KIND == kGyroscope ? int(GyrRec::G0) : int(AccRec::G0));
          constexpr int d1 = (KIND == kCamera) ? int(CamRec::w0) : (KIND == kGyroscope ? int(GyrRec::w0) : int(AccRec::w0));
          constexpr int ji = (KIND == kCamera) ? int(CamRec::Ji) : (KIND == kGyroscope ? int(GyrRec::Ji) : int(AccRec::Ji));
          for (int f = 0; f < m; ++f) rc.put(f, rs * rc.get(f));
          for (int f = d0; f < d1; ++f) rc.put(f, rs * rc.get(f));
          for (int f = ji; f < ji + m * S.ni; ++f) rc.put(f, rs * rc.get(f));
        }
      }
    } else {
      bad = 1;
      if (MODE == kModeJacobian) { const int nf = rec_size(KIND, S.ni); for (int f = 0; f < nf; ++f) rc.put(f, 0.0); }   // rows of zeros
    }
    if (MODE == kModeJacobian) for (int q = 0; q < m; ++q) sd.r[o * m + q] = rc.get(q);
    if (MODE == kModeResiduals) {
      for (int q = 0; q < m; ++q) sd.r[o * m + q] = ok ? rc.get(q) : 0.0;
      // bit 0: the functor evaluated; bit 1 (cameras): the point lies behind the image plane, p_c.z <= 0 — Camera::Project skips those
      // (camera.cpp:172-174,186-188) even for models that can project them.
      sd.valid[o] = ok ? (pc_z > 0.0 ? 1 : 3) : 0;
    }
  }
  s_ok[t] = ok ? 1 : 0;
  // Deterministic block reduction of the cost and of the failure count.
  for (int off = 16; off > 0; off >>= 1) {
    cost += __shfl_down_sync(0xffffffffu, cost, off);
    bad += __shfl_down_sync(0xffffffffu, bad, off);
  }
  if ((t & 31) == 0) { s_red[t >> 5] = cost; s_bad[t >> 5] = bad; }
  __syncthreads();
  if (t == 0) {
    double c = 0.0; int b = 0;
    for (int w = 0; w < kTile / 32; ++w) { c += s_red[w]; b += s_bad[w]; }
    cost_partial[blockIdx.x] = c;
    invalid_partial[blockIdx.x] = b;
  }
  if (MODE == kModeJacobian) {
    // Phase 2. Every Jacobian entry is sum_q rec[fa_q] * rec[fb_q] with (fa, fb) depending only on the position inside the
    // m x jw row block. Each warp streams the CONTIGUOUS region holding the row blocks of its 32 observations, one observation
    // (m * jw doubles) at a time; a lane always produces the same entries of a row block, so their field offsets live in registers.
    constexpr int NQ = (KIND == kCamera) ? 1 : (KIND == kGyroscope ? 2 : 3);
    const int jw = sd.jw, ni = sd.ni;
    const int rowlen = m * jw;
    __syncthreads();
    const int warp = t >> 5, lane = t & 31;
    const int nobs = min(32, tl.count - warp * 32);
    if (nobs > 0) {
      double* __restrict__ Jw = sd.J + (size_t(tl.start) + warp * 32) * rowlen;
      const double* __restrict__ rb = rec + warp * 32;
      if ((rowlen & 1) == 0) {
        // Even row blocks (camera, gyroscope): lane l owns the fixed entry pairs l, l + 32, ... of the row block; the record fields
        // of its entries are loop-invariant registers, so an observation costs the lane 4 NQ shared loads per pair and one 16-byte
        // store; the warp writes each 8 * rowlen-byte row block as contiguous 512-byte pieces. kUnr observations in flight.
        constexpr int P = (KIND == kCamera) ? 2 : 3;               // ceil(m (36 + max stored calibration columns) / 64)
        constexpr int kUnr = 4;
        const int npairs = rowlen >> 1;
        int fa0[P][NQ], fb0[P][NQ], fa1[P][NQ], fb1[P][NQ];
#pragma unroll
        for (int pp = 0; pp < P; ++pp) {
          const int pr = min(lane + 32 * pp, npairs - 1);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int idx = 2 * pr + h;
            const int row = idx / jw, j = idx - row * jw;
            const int canon = j < kCpCols ? j : kCpCols + sd.jcanon[j - kCpCols];
            int fa[3], fb[3];
            jac_terms(KIND, ni, row, canon, fa, fb);
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
              if (h == 0) { fa0[pp][q] = fa[q] * kRecStride; fb0[pp][q] = fb[q] * kRecStride; }
              else { fa1[pp][q] = fa[q] * kRecStride; fb1[pp][q] = fb[q] * kRecStride; }
            }
          }
        }
        for (int o0 = 0; o0 < nobs; o0 += kUnr) {
          double2 v[kUnr][P];
#pragma unroll
          for (int u = 0; u < kUnr; ++u) {
            const double* __restrict__ ro = rb + min(o0 + u, nobs - 1);
#pragma unroll
            for (int pp = 0; pp < P; ++pp) {
              double v0 = 0.0, v1 = 0.0;
#pragma unroll
              for (int q = 0; q < NQ; ++q) { v0 += ro[fa0[pp][q]] * ro[fb0[pp][q]]; v1 += ro[fa1[pp][q]] * ro[fb1[pp][q]]; }
              v[u][pp].x = v0; v[u][pp].y = v1;
            }
          }
#pragma unroll
          for (int u = 0; u < kUnr; ++u)
            if (o0 + u < nobs) {
              double* __restrict__ Jo = Jw + size_t(o0 + u) * rowlen;
#pragma unroll
              for (int pp = 0; pp < P; ++pp)
                if (lane + 32 * pp < npairs) *reinterpret_cast<double2*>(Jo + 2 * (lane + 32 * pp)) = v[u][pp];
            }
        }
      } else {
        // Odd row blocks (accelerometer with an odd column count): same scheme with single entries and 8-byte stores.
        constexpr int P1 = (3 * (kCpCols + kMaxCalib) + 31) / 32;
        constexpr int kUnr = 2;
        int fa1[P1][NQ], fb1[P1][NQ];
#pragma unroll
        for (int pp = 0; pp < P1; ++pp) {
          const int idx = min(lane + 32 * pp, rowlen - 1);
          const int row = idx / jw, j = idx - row * jw;
          const int canon = j < kCpCols ? j : kCpCols + sd.jcanon[j - kCpCols];
          int fa[3], fb[3];
          jac_terms(KIND, ni, row, canon, fa, fb);
#pragma unroll
          for (int q = 0; q < NQ; ++q) { fa1[pp][q] = fa[q] * kRecStride; fb1[pp][q] = fb[q] * kRecStride; }
        }
        for (int o0 = 0; o0 < nobs; o0 += kUnr) {
          double v[kUnr][P1];
#pragma unroll
          for (int u = 0; u < kUnr; ++u) {
            const double* __restrict__ ro = rb + min(o0 + u, nobs - 1);
#pragma unroll
            for (int pp = 0; pp < P1; ++pp) {
              double acc = 0.0;
#pragma unroll
              for (int q = 0; q < NQ; ++q) acc += ro[fa1[pp][q]] * ro[fb1[pp][q]];
              v[u][pp] = acc;
            }
          }
#pragma unroll
          for (int u = 0; u < kUnr; ++u)
            if (o0 + u < nobs) {
              double* __restrict__ Jo = Jw + size_t(o0 + u) * rowlen;
#pragma unroll
              for (int pp = 0; pp < P1; ++pp)
                if (lane + 32 * pp < rowlen) Jo[lane + 32 * pp] = v[u][pp];
            }
        }
      }
    }
  }
}

// K0: one thread per camera image (sensor, stamp): basis weights, spline pose and velocity at stamp - latency, R_rw, J_l(phi),
// R_rc^T -> frames[f][FrameRec::kSize]. The 25..144 residual blocks of the image read it back through L1 (warp-broadcast).
__global__ void __launch_bounds__(32) camera_frame_kernel(int n_frames, const int* __restrict__ frame_sensor, const int* __restrict__ frame_seg,
                                                           const double* __restrict__ frame_stamp, const SensorState* __restrict__ states,
                                                           const double* __restrict__ ctrl, const double* __restrict__ knots,
                                                           const double* __restrict__ basis, double* __restrict__ frames) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_frames) return;
  const int seg = frame_seg[f];
  double fr[FrameRec::kSize];
  camera_frame(states[frame_sensor[f]], basis + size_t(seg) * (kK * kK), knots[seg + kK - 1], knots[seg + kK], ctrl + size_t(seg) * 6, frame_stamp[f], fr);
  double* out = frames + size_t(f) * FrameRec::kSize;
#pragma unroll
  for (int i = 0; i < FrameRec::kSize; ++i) out[i] = fr[i];
}

// Sums the per-tile partials in a fixed order: scal[slot] = cost, scal[slot+1] = number of failed blocks. One CTA of 1024 threads,
// 4 independent loads in flight per thread, shuffle reduction.
__global__ void __launch_bounds__(1024) reduce_cost_kernel(const double* __restrict__ cost_partial, const int* __restrict__ invalid_partial,
                                                           int n, double* __restrict__ scal, int slot) {
  __shared__ double sc[32];
  __shared__ int sb[32];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  double c = 0.0; int b = 0;
  for (int i0 = t; i0 < n; i0 += 4 * 1024) {
    double cv[4]; int bv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int i = i0 + u * 1024; const int ic = min(i, n - 1); cv[u] = cost_partial[ic]; bv[u] = invalid_partial[ic]; if (i >= n) { cv[u] = 0.0; bv[u] = 0; } }
#pragma unroll
    for (int u = 0; u < 4; ++u) { c += cv[u]; b += bv[u]; }
  }
  for (int off = 16; off > 0; off >>= 1) { c += __shfl_down_sync(0xffffffffu, c, off); b += __shfl_down_sync(0xffffffffu, b, off); }
  if (lane == 0) { sc[warp] = c; sb[warp] = b; }
  __syncthreads();
  if (warp == 0) {
    c = sc[lane]; b = sb[lane];
    for (int off = 16; off > 0; off >>= 1) { c += __shfl_down_sync(0xffffffffu, c, off); b += __shfl_down_sync(0xffffffffu, b, off); }
    if (lane == 0) { scal[slot] = c; scal[slot + 1] = double(b); }
  }
}

}  // namespace cb2
