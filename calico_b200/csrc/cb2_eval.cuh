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
  // fused compact Gram (cameras, Jacobian mode): image groups of this tile
  __shared__ int s_gstart[(KIND == kCamera && MODE == kModeJacobian) ? kTile + 1 : 1], s_gfrm[(KIND == kCamera && MODE == kModeJacobian) ? kTile : 1];
  __shared__ int s_wcount[kTile / 32], s_ng, s_hole;
  constexpr int m = (KIND == kCamera) ? 2 : 3;
  const EvalTile tl = tiles[blockIdx.x];
  const SensorDesc& sd = sensors[tl.sensor];
  const int t = threadIdx.x;
  const bool active = t < tl.count;
  double cost = 0.0;
  int bad = 0;
  bool ok = false;
  double pc_z = 1.0;   // cameras, kModeResiduals: depth of the point in the camera frame (flag 2 of `valid`)
  constexpr bool kGram = KIND == kCamera && MODE == kModeJacobian;
  const bool do_gram = kGram && sd.gslots != nullptr;
  int my_frm = -1, prev_frm = -2;
  unsigned start_mask = 0;
  bool is_start = false;
  if (kGram && do_gram) {
    // Image (frame) of every observation of the tile; an observation starts a group when its image differs from its predecessor's.
    if (active) my_frm = sd.frm[long(tl.start) + t];
    prev_frm = __shfl_up_sync(0xffffffffu, my_frm, 1);
    if ((t & 31) == 0) prev_frm = (active && long(tl.start) + t > 0) ? sd.frm[long(tl.start) + t - 1] : -2;
    is_start = active && (t == 0 || my_frm != prev_frm);
    start_mask = __ballot_sync(0xffffffffu, is_start);
    if ((t & 31) == 0) s_wcount[t >> 5] = __popc(start_mask);
    if (t == 0) s_hole = (tl.start > 0 && my_frm != prev_frm) ? 1 : 0;   // the tile begins exactly at an image boundary: slot (first - 1) stays unused
  }
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
  if (kGram && do_gram) {
    int base = 0;
    for (int w = 0; w < (t >> 5); ++w) base += s_wcount[w];
    if (is_start) { const int gi = base + __popc(start_mask & ((1u << (t & 31)) - 1u)); s_gstart[gi] = t; s_gfrm[gi] = my_frm; }
    if (t == 0) { int ng = 0; for (int w = 0; w < kTile / 32; ++w) ng += s_wcount[w]; s_ng = ng; }
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
    // Compact Gram of every image group of the tile on the FP64 tensor pipe, straight from the records in shared memory: rows are the
    // residual rows (observation o, q in {0, 1}), columns [g_0..g_5 | r | 0 | calibration unknowns 0..15]. One k-step = 4 rows = 2
    // observations; lane l supplies row (o + (l >> 1 & 1), q = l & 1) of column 8 b + (l >> 2) — a fixed record field per column block.
    // The tiles that need the image's basis weights go to the image's slot; the calibration tiles are summed over the warp's groups.
    double cgS[5][2];
#pragma unroll
    for (int i = 0; i < 5; ++i) { cgS[i][0] = 0.0; cgS[i][1] = 0.0; }
    auto gram_phase = [&]() {
      const int q = lane & 1, oo = (lane >> 1) & 1, c = lane >> 2;
      int fo[3];
#pragma unroll
      for (int b = 0; b < 3; ++b) fo[b] = int(sd.gfield[q][8 * b + c]) * kRecStride;
      const int ng = s_ng;
      const int cta_in_sensor = tl.start / kTile;
      const int ri = lane >> 2, jp = lane & 3;          // accumulator fragment: row ri, columns 2 jp, 2 jp + 1
      for (int g = warp; g < ng; g += kTile / 32) {
        const int a = s_gstart[g], e = g + 1 < ng ? s_gstart[g + 1] : tl.count;
        double cg[6][2];
#pragma unroll
        for (int i = 0; i < 6; ++i) { cg[i][0] = 0.0; cg[i][1] = 0.0; }
        const double* ro = rec + a + oo;
        const int npair = (e - a) >> 1;
#pragma unroll 2
        for (int kp = 0; kp < npair; ++kp, ro += 2) {
          const double x0 = ro[fo[0]], x1 = ro[fo[1]], x2 = ro[fo[2]];
          dmma_8x8x4(cg[0][0], cg[0][1], x0, x0);
          dmma_8x8x4(cg[1][0], cg[1][1], x1, x0);
          dmma_8x8x4(cg[2][0], cg[2][1], x1, x1);
          dmma_8x8x4(cg[3][0], cg[3][1], x2, x0);
          dmma_8x8x4(cg[4][0], cg[4][1], x2, x1);
          dmma_8x8x4(cg[5][0], cg[5][1], x2, x2);
        }
        if ((e - a) & 1) {                              // odd group: the last k-step holds one observation (rows of lanes with oo == 1 are zero)
          const double* rl = rec + (e - 1);
          const double x0 = oo ? 0.0 : rl[fo[0]], x1 = oo ? 0.0 : rl[fo[1]], x2 = oo ? 0.0 : rl[fo[2]];
          dmma_8x8x4(cg[0][0], cg[0][1], x0, x0);
          dmma_8x8x4(cg[1][0], cg[1][1], x1, x0);
          dmma_8x8x4(cg[2][0], cg[2][1], x1, x1);
          dmma_8x8x4(cg[3][0], cg[3][1], x2, x0);
          dmma_8x8x4(cg[4][0], cg[4][1], x2, x1);
          dmma_8x8x4(cg[5][0], cg[5][1], x2, x2);
        }
        const int slot = sd.gslot_base + s_gfrm[g] + cta_in_sensor;
        if (jp < 3) {                                   // columns 0..5 (g) of the tiles (0,0), (1,0), (2,0): rows of 6 doubles
          double* __restrict__ S = sd.gslots + size_t(slot) * kGramSlot + ri * 6 + 2 * jp;
          *reinterpret_cast<double2*>(S) = make_double2(cg[0][0], cg[0][1]);
          *reinterpret_cast<double2*>(S + 48) = make_double2(cg[1][0], cg[1][1]);
          *reinterpret_cast<double2*>(S + 96) = make_double2(cg[3][0], cg[3][1]);
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) { cgS[i][0] += cg[i + 1][0]; cgS[i][1] += cg[i + 1][1]; }
        {
          double* __restrict__ S = sd.gslots + size_t(slot) * kGramSlot;
          if (lane < kK) S[kGramSlotW + lane] = rec[(CamRec::w0 + lane) * kRecStride + a];   // the image's basis weights (same for all its rows)
          else if (lane == kK) S[kGramSlotFlag] = 1.0;
          else if (lane == kK + 1 && g == 0 && s_hole) S[kGramSlotFlag - kGramSlot] = 0.0;    // the tile starts at an image boundary: slot - 1 is unused
        }
      }
    };
#ifdef CB2_GRAM_FIRST
    if (kGram && do_gram) gram_phase();
#endif
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
    if (kGram && do_gram) {
#ifndef CB2_GRAM_FIRST
      gram_phase();                                     // the Jacobian stores of this warp drain while its Gram tiles are multiplied
#endif
      // Calibration tiles summed over this WARP's groups -> gcta[CTA][warp]: no block barrier, no atomics; assemble_calib_kernel sums the
      // (CTA, warp) partials in a fixed order. Layout: [calib 0..7 x calib 0..7 | calib 8..15 x 0..7 | 8..15 x 8..15 | gradient 16].
      double* __restrict__ out = sd.gcta + (size_t(tl.start / kTile) * (kTile / 32) + warp) * kGramCta;
      *reinterpret_cast<double2*>(out + 2 * lane) = make_double2(cgS[1][0], cgS[1][1]);
      *reinterpret_cast<double2*>(out + 64 + 2 * lane) = make_double2(cgS[3][0], cgS[3][1]);
      *reinterpret_cast<double2*>(out + 128 + 2 * lane) = make_double2(cgS[4][0], cgS[4][1]);
      if ((lane & 3) == 3) { out[192 + (lane >> 2)] = cgS[0][0]; out[200 + (lane >> 2)] = cgS[2][0]; }   // column 6 (r) of calib x [g | r]
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
#ifdef CB2_EMUL
constexpr int kRcThreads = 128;    // (emulation build: fewer OS threads)
#else
constexpr int kRcThreads = 1024;
#endif
__global__ void __launch_bounds__(kRcThreads) reduce_cost_kernel(const double* __restrict__ cost_partial, const int* __restrict__ invalid_partial,
                                                           int n, double* __restrict__ scal, int slot) {
  __shared__ double sc[32];
  __shared__ int sb[32];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  double c = 0.0; int b = 0;
  for (int i0 = t; i0 < n; i0 += 4 * kRcThreads) {
    double cv[4]; int bv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int i = i0 + u * kRcThreads; const int ic = min(i, n - 1); cv[u] = cost_partial[ic]; bv[u] = invalid_partial[ic]; if (i >= n) { cv[u] = 0.0; bv[u] = 0; } }
#pragma unroll
    for (int u = 0; u < 4; ++u) { c += cv[u]; b += bv[u]; }
  }
  for (int off = 16; off > 0; off >>= 1) { c += __shfl_down_sync(0xffffffffu, c, off); b += __shfl_down_sync(0xffffffffu, b, off); }
  if (lane == 0) { sc[warp] = c; sb[warp] = b; }
  __syncthreads();
  if (warp == 0) {
    c = lane < kRcThreads / 32 ? sc[lane] : 0.0; b = lane < kRcThreads / 32 ? sb[lane] : 0;
    for (int off = 16; off > 0; off >>= 1) { c += __shfl_down_sync(0xffffffffu, c, off); b += __shfl_down_sync(0xffffffffu, b, off); }
    if (lane == 0) { scal[slot] = c; scal[slot + 1] = double(b); }
  }
}

}  // namespace cb2
