// calico_b200 — small per-iteration kernels of the Levenberg-Marquardt loop (Ceres external semantics:
// trust_region_minimizer.cc, levenberg_marquardt_strategy.cc, manifold.h), all single-CTA with fixed reduction order.
#pragma once
#include "cb2_device.cuh"

namespace cb2 {

#ifdef CB2_EMUL
constexpr int kLmThreads = 128;    // emulation build: one OS thread per CUDA thread, the barrier-heavy reductions get 8x fewer of them
#else
constexpr int kLmThreads = 1024;
#endif

// diag[j] = H(j, j): control points from the band, calibration from C.
__global__ void __launch_bounds__(256) hess_diag_kernel(long n_a, int N_c, const double* __restrict__ Aband, const double* __restrict__ Cmat,
                                                        double* __restrict__ diag) {
  const long n = n_a + N_c;
  for (long j = long(blockIdx.x) * blockDim.x + threadIdx.x; j < n; j += long(gridDim.x) * blockDim.x)
    diag[j] = j < n_a ? Aband[j * kCpCols + (kCpCols - 1)] : Cmat[(j - n_a) * N_c + (j - n_a)];
}

// Jacobi scaling fixed at iteration 0: s_j = 1 / (1 + sqrt(diag_j))  (TrustRegionMinimizer::EvaluateGradientAndJacobian).
__global__ void __launch_bounds__(256) jacobi_scaling_kernel(long n, const double* __restrict__ diag, int enable, double* __restrict__ scaling) {
  for (long j = long(blockIdx.x) * blockDim.x + threadIdx.x; j < n; j += long(gridDim.x) * blockDim.x)
    scaling[j] = enable ? 1.0 / (1.0 + sqrt(diag[j])) : 1.0;
}

// LM damping in unscaled coordinates. Ceres solves (S H S + D^2) y = S g with D^2 = clamp(diag(S H S), lo, hi) / radius and
// applies delta = -S y. With ytil = S y this is (H + Dtil^2) ytil = g, Dtil^2_j = clamp(s_j^2 H_jj, lo, hi) / (radius s_j^2).
__global__ void __launch_bounds__(256) damping_kernel(long n, const double* __restrict__ diag, const double* __restrict__ scaling,
                                                      const double* __restrict__ scal, double* __restrict__ dtil2) {
  // radius and the clamp bounds come from the scalar block (written by the host right before the launch), not from kernel arguments,
  // so that the whole solve phase can be replayed as one CUDA graph with a different trust-region radius every iteration.
  const double radius = scal[kScRadius], lo = scal[kScLmLo], hi = scal[kScLmHi];
  for (long j = long(blockIdx.x) * blockDim.x + threadIdx.x; j < n; j += long(gridDim.x) * blockDim.x) {
    const double s2 = scaling[j] * scaling[j];
    dtil2[j] = fmin(fmax(diag[j] * s2, lo), hi) / (radius * s2);
  }
}

// Fixed-order block reductions of NV values at once (all kLmThreads threads must call): warp shuffles, then the first warp combines the
// per-warp partials. sh must hold NV * 32 doubles. The result is valid in thread 0.
template <int NV, bool kMax>
CB2_D void block_reduce(double (&v)[NV], double* sh) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int q = 0; q < NV; ++q) { const double o = __shfl_down_sync(0xffffffffu, v[q], off); v[q] = kMax ? fmax(v[q], o) : v[q] + o; }
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int q = 0; q < NV; ++q) sh[q * 32 + warp] = v[q];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) v[q] = lane < kLmThreads / 32 ? sh[q * 32 + lane] : (kMax ? -1.0e308 : 0.0);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
      for (int q = 0; q < NV; ++q) { const double o = __shfl_down_sync(0xffffffffu, v[q], off); v[q] = kMax ? fmax(v[q], o) : v[q] + o; }
  }
}

// Grid-wide finish of a block_reduce: thread 0 of every CTA leaves its NV values in partial[blockIdx.x][NV]; the LAST CTA to arrive
// (atomic ticket) gets true in its thread 0 with v[] = the CTAs' values combined in CTA order (fixed order -> deterministic), and resets
// the ticket for the next launch / graph replay. max_mask bit q: value q is combined by maximum instead of sum.
template <int NV>
CB2_D bool grid_reduce_last(double (&v)[NV], unsigned max_mask, double* __restrict__ partial, unsigned* __restrict__ ticket) {
  if (threadIdx.x != 0) return false;
  if (gridDim.x == 1) return true;
#pragma unroll
  for (int q = 0; q < NV; ++q) partial[blockIdx.x * NV + q] = v[q];
  __threadfence();
  if (atomicAdd(ticket, 1u) != gridDim.x - 1) return false;
  __threadfence();
  const volatile double* pv = partial;
#pragma unroll
  for (int q = 0; q < NV; ++q) v[q] = pv[q];
  for (unsigned b = 1; b < gridDim.x; ++b)
#pragma unroll
    for (int q = 0; q < NV; ++q) { const double o = pv[b * NV + q]; v[q] = (max_mask >> q & 1u) ? fmax(v[q], o) : v[q] + o; }
  *ticket = 0u;
  return true;
}
#ifdef CB2_EMUL
constexpr int kLmMaxCtas = 2;       // emulation build: two CTAs even on the micro problems of the CPU tests (any grid size is valid)
constexpr long kLmQuantum = 256;   // (micro / tiny problems: one CTA; the `small` ones: two)
#else
constexpr int kLmMaxCtas = 32;      // CTAs of gradient_norm_kernel / apply_step_kernel (4 x kLmThreads unknowns each)
constexpr long kLmQuantum = 4L * kLmThreads;
#endif
CB2_HD int lm_ctas(long n_a) { const long c = (n_a + kLmQuantum - 1) / kLmQuantum; return int(c < 1 ? 1 : (c > kLmMaxCtas ? kLmMaxCtas : c)); }

// Freed world-model blocks (see world_points_kernel below): where the rigid-body poses and model points sit in the calibration vector.
struct WorldRefs {
  int n_bodies, n_points;
  const int* body_u;   // [2 n_bodies] offsets of (rotation, translation) in the calibration vector, -1 = constant
  const int* pt_u;     // [n_points] offset of the model point, -1 = constant
};

// gradient_max_norm / gradient_norm^2 = |x - Plus(x, -g)|_inf / _2^2 (ambient coordinates) over a part of the reduced parameter vector:
// count_owned: the control points this rank owns; count_shared: the separator control points and the calibration blocks (whose gradient
// entries are sums over the ranks). combine: fold the result into what scal[kScGradSq / kScGradMax] already hold (multi-rank: the
// cross-rank sum / maximum of the ranks' owned parts) instead of overwriting them.
__global__ void __launch_bounds__(kLmThreads) gradient_norm_kernel(long n_a, const double* __restrict__ grad, const unsigned char* __restrict__ cp_own,
                                                                   int count_owned, int count_shared, int combine, const SensorDesc* __restrict__ sensors,
                                                                   const SensorState* __restrict__ states, int n_sensors, WorldRefs world,
                                                                   const double* __restrict__ body_q, double* __restrict__ scal,
                                                                   double* __restrict__ partial, unsigned* __restrict__ ticket) {
  __shared__ double sh[32];
  const int t = threadIdx.x;
  double mx = 0.0, sq = 0.0;
  // 4 independent loads in flight per thread; lm_ctas(n_a) CTAs of 4 kLmThreads unknowns each (the kernel is pure latency: one round per CTA)
  for (long i0 = blockIdx.x * 4L * kLmThreads + t; i0 < n_a; i0 += gridDim.x * 4L * kLmThreads) {
    double g[4]; int own[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const long i = min(i0 + long(u) * kLmThreads, n_a - 1); g[u] = grad[i]; own[u] = cp_own[i / 6]; }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i0 + long(u) * kLmThreads < n_a && ((own[u] == kCpOwned && count_owned) || (own[u] == kCpShared && count_shared))) { mx = fmax(mx, fabs(g[u])); sq += g[u] * g[u]; }
  }
  const bool cal = count_shared && blockIdx.x == 0;       // the calibration blocks: first CTA only
  if (cal) for (int s = t; s < n_sensors; s += kLmThreads) {
    const SensorDesc& sd = sensors[s];
    const double* gc = grad + n_a + sd.calib_off;
    for (int j = 0; j < sd.n_calib; ++j) {
      if (sd.u_rot >= 0 && j >= sd.u_rot && j < sd.u_rot + 3) continue;
      mx = fmax(mx, fabs(gc[j])); sq += gc[j] * gc[j];
    }
    if (sd.u_rot >= 0) {
      const Q4 q = states[s].q;
      const Q4 p = quat_plus(q, v3(-gc[sd.u_rot], -gc[sd.u_rot + 1], -gc[sd.u_rot + 2]));
      const double d[4] = {q.x - p.x, q.y - p.y, q.z - p.z, q.w - p.w};
      for (int k = 0; k < 4; ++k) { mx = fmax(mx, fabs(d[k])); sq += d[k] * d[k]; }
    }
  }
  if (cal) {   // freed world-model blocks: pose rotation through the manifold, the rest directly
    for (int b = t; b < world.n_bodies; b += kLmThreads) {
      const int ur = world.body_u[2 * b], ut = world.body_u[2 * b + 1];
      if (ur >= 0) {
        const double* gc = grad + n_a + ur;
        const Q4 q = Q4{body_q[4 * b], body_q[4 * b + 1], body_q[4 * b + 2], body_q[4 * b + 3]};
        const Q4 p = quat_plus(q, v3(-gc[0], -gc[1], -gc[2]));
        const double d[4] = {q.x - p.x, q.y - p.y, q.z - p.z, q.w - p.w};
        for (int k = 0; k < 4; ++k) { mx = fmax(mx, fabs(d[k])); sq += d[k] * d[k]; }
      }
      if (ut >= 0) for (int k = 0; k < 3; ++k) { const double g = grad[n_a + ut + k]; mx = fmax(mx, fabs(g)); sq += g * g; }
    }
    for (int p = t; p < world.n_points; p += kLmThreads) {
      const int u = world.pt_u[p];
      if (u >= 0) for (int k = 0; k < 3; ++k) { const double g = grad[n_a + u + k]; mx = fmax(mx, fabs(g)); sq += g * g; }
    }
  }
  double vs[1] = {sq}, vm[1] = {mx};
  block_reduce<1, false>(vs, sh);
  block_reduce<1, true>(vm, sh);
  double both[2] = {vs[0], vm[0]};
  if (grid_reduce_last<2>(both, 2u, partial, ticket)) {
    if (combine) { both[1] = fmax(both[1], scal[kScGradMax]); both[0] += scal[kScGradSq]; }
    scal[kScGradMax] = both[1]; scal[kScGradSq] = both[0];
  }
}

// Candidate point x_cand = Plus(x, -ytil) (owned + shared control points and every non-constant sensor block), with the part
// this rank counts of
//   step_norm^2 = |x - x_cand|^2, x_norm^2 = |x|^2, cand_x_norm^2 (ambient, reduced program only: cp_ref marks control
//   points referenced by at least one residual block), the model cost change 1/2 ytil.(g + Dtil^2 ytil)
//   (== -(J step)^T (r + J step / 2) for the exact solution of the damped system) and a finiteness check of the step.
__global__ void __launch_bounds__(kLmThreads) apply_step_kernel(long n_a, const double* __restrict__ ytil, const double* __restrict__ grad,
                                                                const double* __restrict__ dtil2, const unsigned char* __restrict__ cp_ref,
                                                                const unsigned char* __restrict__ cp_own, int count_shared,
                                                                const double* __restrict__ ctrl, double* __restrict__ ctrl_cand,
                                                                const SensorDesc* __restrict__ sensors, const SensorState* __restrict__ states,
                                                                SensorState* __restrict__ states_cand, int n_sensors, int N_c,
                                                                WorldRefs world, const double* __restrict__ body_q, const double* __restrict__ body_t,
                                                                const double* __restrict__ pm, double* __restrict__ body_q_cand,
                                                                double* __restrict__ body_t_cand, double* __restrict__ pm_cand, double* __restrict__ scal,
                                                                double* __restrict__ partial, unsigned* __restrict__ ticket) {
  __shared__ double sh[5 * 32];
  const int t = threadIdx.x;
  double step2 = 0.0, x2 = 0.0, c2 = 0.0, model = 0.0, bad = 0.0;
  // 4 elements per thread per round, all loads issued before their first use; lm_ctas(n_a) CTAs, the calibration / world blocks in the first
  for (long i0 = blockIdx.x * 4L * kLmThreads + t; i0 < n_a; i0 += gridDim.x * 4L * kLmThreads) {
    double xv[4], yv[4], gv[4], dv[4]; int own[4], ref[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long i = min(i0 + long(u) * kLmThreads, n_a - 1);
      xv[u] = ctrl[i]; yv[u] = ytil[i]; gv[u] = grad[i]; dv[u] = dtil2[i]; own[u] = cp_own[i / 6]; ref[u] = cp_ref[i / 6];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long i = i0 + long(u) * kLmThreads;
      if (i >= n_a) continue;
      const double x = xv[u];
      if (own[u] == kCpPeer) { ctrl_cand[i] = x; continue; }
      const double y = yv[u];
      const double xn = x - y;
      ctrl_cand[i] = xn;
      if (!isfinite(y)) bad = 1.0;
      if (own[u] == kCpOwned || count_shared) {
        model += y * (gv[u] + dv[u] * y);
        if (ref[u]) { step2 += (x - xn) * (x - xn); x2 += x * x; c2 += xn * xn; }
      }
    }
  }
  const bool first = blockIdx.x == 0;
  for (long j = t; first && j < N_c; j += kLmThreads) {
    const double y = ytil[n_a + j];
    if (!isfinite(y)) bad = 1.0;
    if (count_shared) model += y * (grad[n_a + j] + dtil2[n_a + j] * y);
  }
  const double cs = count_shared ? 1.0 : 0.0;
  for (int s = t; first && s < n_sensors; s += kLmThreads) {
    const SensorDesc& sd = sensors[s];
    SensorState S = states[s];
    const double* y = ytil + n_a + sd.calib_off;
    if (sd.u_intr >= 0) for (int j = 0; j < sd.ni; ++j) {
      const double x = S.intr[j], xn = x - y[sd.u_intr + j];
      S.intr[j] = xn; step2 += cs * (x - xn) * (x - xn); x2 += cs * x * x; c2 += cs * xn * xn;
    }
    if (sd.u_rot >= 0) {
      const Q4 q = S.q;
      const Q4 p = quat_plus(q, v3(-y[sd.u_rot], -y[sd.u_rot + 1], -y[sd.u_rot + 2]));
      S.q = p;
      step2 += cs * ((q.x - p.x) * (q.x - p.x) + (q.y - p.y) * (q.y - p.y) + (q.z - p.z) * (q.z - p.z) + (q.w - p.w) * (q.w - p.w));
      x2 += cs * (q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
      c2 += cs * (p.x * p.x + p.y * p.y + p.z * p.z + p.w * p.w);
    }
    if (sd.u_trans >= 0) {
      const V3 x = S.t;
      const V3 xn = v3(x.x - y[sd.u_trans], x.y - y[sd.u_trans + 1], x.z - y[sd.u_trans + 2]);
      S.t = xn; step2 += cs * dot(x - xn, x - xn); x2 += cs * dot(x, x); c2 += cs * dot(xn, xn);
    }
    if (sd.u_lat >= 0) {
      const double x = S.latency, xn = x - y[sd.u_lat];
      S.latency = xn; step2 += cs * (x - xn) * (x - xn); x2 += cs * x * x; c2 += cs * xn * xn;
    }
    states_cand[s] = S;
  }
  // freed world-model blocks (constant ones are copied so that the candidate tables are complete)
  for (int b = t; first && b < world.n_bodies; b += kLmThreads) {
    const int ur = world.body_u[2 * b], ut = world.body_u[2 * b + 1];
    const Q4 q = Q4{body_q[4 * b], body_q[4 * b + 1], body_q[4 * b + 2], body_q[4 * b + 3]};
    Q4 p = q;
    if (ur >= 0) {
      const double* y = ytil + n_a + ur;
      p = quat_plus(q, v3(-y[0], -y[1], -y[2]));
      step2 += cs * ((q.x - p.x) * (q.x - p.x) + (q.y - p.y) * (q.y - p.y) + (q.z - p.z) * (q.z - p.z) + (q.w - p.w) * (q.w - p.w));
      x2 += cs * (q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
      c2 += cs * (p.x * p.x + p.y * p.y + p.z * p.z + p.w * p.w);
    }
    body_q_cand[4 * b] = p.x; body_q_cand[4 * b + 1] = p.y; body_q_cand[4 * b + 2] = p.z; body_q_cand[4 * b + 3] = p.w;
    for (int k = 0; k < 3; ++k) {
      const double x = body_t[3 * b + k];
      double xn = x;
      if (ut >= 0) { xn = x - ytil[n_a + ut + k]; step2 += cs * (x - xn) * (x - xn); x2 += cs * x * x; c2 += cs * xn * xn; }
      body_t_cand[3 * b + k] = xn;
    }
  }
  for (int p = t; first && p < world.n_points; p += kLmThreads) {
    const int u = world.pt_u[p];
    for (int k = 0; k < 3; ++k) {
      const double x = pm[3 * p + k];
      double xn = x;
      if (u >= 0) { xn = x - ytil[n_a + u + k]; step2 += cs * (x - xn) * (x - xn); x2 += cs * x * x; c2 += cs * xn * xn; }
      pm_cand[3 * p + k] = xn;
    }
  }
  double red[5] = {step2, x2, c2, model, bad};
  block_reduce<5, false>(red, sh);
  if (grid_reduce_last<5>(red, 0u, partial, ticket)) {
    scal[kScStepNorm2] = red[0]; scal[kScXNorm2] = red[1]; scal[kScCandXNorm2] = red[2]; scal[kScModelChange] = 0.5 * red[3];
    if (red[4] > 0.0) scal[kScSolveFail] += 1.0;
  }
}

// Multi-rank glue. Separator rows and calibration receive contributions from several ranks: their gradient and Hessian-diagonal entries
// (needed for the gradient norms, the Jacobi scaling and the LM damping), the cost / failure count of the sweep and the ranks' owned parts
// of the gradient norms travel in ONE buffer through one cross-rank sum:
//   buf = [grad(idx[0..n)) | diag(idx[0..n)) | cost, invalid, owned |g|^2 | one slot per rank holding that rank's owned |g|_inf, zero elsewhere].
// idx[i] = global unknown index of shared entry i. After the sum every rank takes the maximum of the rank slots.
CB2_HD constexpr size_t shared_buf_size(int n_shared, int world) { return 2 * size_t(n_shared) + 3 + size_t(world); }
__global__ void __launch_bounds__(256) pack_shared_kernel(int n, const int* __restrict__ idx, const double* __restrict__ grad, const double* __restrict__ diag,
                                                          const double* __restrict__ scal, int world, int rank, double* __restrict__ buf) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { buf[i] = grad[idx[i]]; buf[n + i] = diag[idx[i]]; }
  if (blockIdx.x == 0) {
    double* tail = buf + 2 * size_t(n);
    const int t = threadIdx.x;
    if (t < 3) tail[t] = scal[kScCost + t];             // kScCost, kScInvalid, kScGradSq are consecutive
    for (int r = t; r < world; r += blockDim.x) tail[3 + r] = r == rank ? scal[kScGradMax] : 0.0;
  }
}
// gradG = this rank's gradient with the summed entries on the shared rows; diag likewise; the scalar sums back into scal.
__global__ void __launch_bounds__(256) unpack_shared_kernel(int n, const int* __restrict__ idx, const double* __restrict__ buf, int world,
                                                            double* __restrict__ gradG, double* __restrict__ diag, double* __restrict__ scal) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { gradG[idx[i]] = buf[i]; diag[idx[i]] = buf[n + i]; }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double* tail = buf + 2 * size_t(n);
    for (int q = 0; q < 3; ++q) scal[kScCost + q] = tail[q];
    double m = 0.0;
    for (int r = 0; r < world; ++r) m = fmax(m, tail[3 + r]);
    scal[kScGradMax] = m;
  }
}
// LM damping of the shared rows once their summed diagonal is known (see damping_kernel).
__global__ void __launch_bounds__(256) damping_shared_kernel(int n, const int* __restrict__ idx, const double* __restrict__ diag, const double* __restrict__ scaling,
                                                             const double* __restrict__ scal, double* __restrict__ dtil2) {
  const double radius = scal[kScRadius], lo = scal[kScLmLo], hi = scal[kScLmHi];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int j = idx[i];
    const double s2 = scaling[j] * scaling[j];
    dtil2[j] = fmin(fmax(diag[j] * s2, lo), hi) / (radius * s2);
  }
}

// Sensor::UpdateResiduals: residuals from the device order (sorted by spline segment, this rank's shard) back into the caller's observation
// order. perm[i] = original index of sorted position i; rfull / vfull were zeroed (outliers and other ranks' observations keep 0).
__global__ void __launch_bounds__(256) scatter_residuals_kernel(int n_active, int m, const int* __restrict__ perm, const double* __restrict__ r,
                                                                const unsigned char* __restrict__ valid, double* __restrict__ rfull,
                                                                unsigned char* __restrict__ vfull) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_active; i += gridDim.x * blockDim.x) {
    const long o = perm[i];
    for (int q = 0; q < m; ++q) rfull[o * m + q] = r[long(i) * m + q];
    vfull[o] = valid[i] ? 1 : 0;
  }
}

// ----------------------------------------------------------------------------------------------------------------
// Freed world-model blocks (world_model.cpp:40-77: RigidBody::world_pose_is_constant / model_definition_is_constant = false). The rigid
// body pose (quaternion with the left-multiplicative half-angle tangent of ceres::EigenQuaternionManifold + translation) and its model
// points become unknowns appended to the calibration vector: body b: u_rot = body_u[2 b], u_trans = body_u[2 b + 1]; world point p:
// pt_u[p]; -1 = constant. World points p_w = R_wm p_m + t_wm are recomputed from the current values (one table per parameter buffer).
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) world_points_kernel(int n_points, const int* __restrict__ pt_body, const double* __restrict__ body_q,
                                                           const double* __restrict__ body_t, const double* __restrict__ pm, double* __restrict__ pw) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_points; p += gridDim.x * blockDim.x) {
    const int b = pt_body[p];
    const M3 R = quat_matrix(Q4{body_q[4 * b], body_q[4 * b + 1], body_q[4 * b + 2], body_q[4 * b + 3]});
    const V3 v = R * v3(pm[3 * p], pm[3 * p + 1], pm[3 * p + 2]);
    pw[3 * p] = v.x + body_t[3 * b]; pw[3 * p + 1] = v.y + body_t[3 * b + 1]; pw[3 * p + 2] = v.z + body_t[3 * b + 2];
  }
}

// Normal-equation contributions of the world unknowns of one camera's residual blocks, from the stored (robustified) Jacobian rows:
//   d r / d p_w = - sum_i J[., 6 i + 3 .. 6 i + 5]   (p_w enters the residual only through p_w - t_wr; the basis weights sum to one), then
//   d r / d t_wm = d r / d p_w,   d r / d delta_wm = 2 (v x d) per row with v = p_w - t_wm (left-multiplicative half-angle tangent),
//   d r / d p_m = (d r / d p_w) R_wm.
// Every product with a world column — against the control points of the block's segment, the camera's calibration columns, the other
// world columns and the residual — is ADDED to the assembled Bmat / Cmat / gradient (after assemble_*_kernel wrote them). The pose columns
// are shared by all observations of a body: a CTA sums them in shared memory (shared-memory atomics) and flushes once; the point columns go
// straight to global atomics. Floating-point atomics make the summation order — and so the last bits of these entries — run-dependent:
// this optional path trades the bitwise repeatability of the rest of the pipeline for simplicity.
constexpr int kWorldMaxSeg = 8;    // spline segments one 128-observation tile may span in the shared-memory path (more: direct atomics)
__global__ void __launch_bounds__(128) world_normal_kernel(const SensorDesc* __restrict__ sensors, const EvalTile* __restrict__ tiles, long n_a, int N_c,
                                                           const int* __restrict__ pt_body, const int* __restrict__ body_u, const int* __restrict__ pt_u,
                                                           const double* __restrict__ body_q, const double* __restrict__ body_t,
                                                           const double* __restrict__ pw, double* __restrict__ Bmat, double* __restrict__ Cmat,
                                                           double* __restrict__ grad) {
  constexpr int kRow = 1 + 6 + kMaxCalib + kWorldMaxSeg * kCpCols;      // per pose column: gradient | pose columns | calibration | cp of up to 8 segments
  __shared__ double s_acc[6 * kRow];
  __shared__ int s_home[2];                                             // body and first segment of the tile's first observation
  const EvalTile tl = tiles[blockIdx.x];
  const SensorDesc& sd = sensors[tl.sensor];
  const int t = threadIdx.x;
  for (int e = t; e < 6 * kRow; e += blockDim.x) s_acc[e] = 0.0;
  if (t == 0) { const long o0 = tl.start; s_home[0] = pt_body[sd.pt[o0]]; s_home[1] = sd.seg[o0]; }
  __syncthreads();
  const int home_b = s_home[0], seg0 = s_home[1];
  const int jw = sd.jw, njc = sd.n_jcal;
  if (t < tl.count) {
    const long o = long(tl.start) + t;
    const int p = sd.pt[o], b = pt_body[p], seg = sd.seg[o];
    int uw[9];
    uw[0] = uw[1] = uw[2] = body_u[2 * b]; uw[3] = uw[4] = uw[5] = body_u[2 * b + 1]; uw[6] = uw[7] = uw[8] = pt_u[p];
#pragma unroll
    for (int a = 0; a < 9; ++a) uw[a] = uw[a] < 0 ? -1 : uw[a] + a % 3;
    if (uw[0] >= 0 || uw[3] >= 0 || uw[6] >= 0) {
      const double* J0 = sd.J + size_t(o) * 2 * jw;
      const double* J1 = J0 + jw;
      const double r0 = sd.r[2 * o], r1 = sd.r[2 * o + 1];
      double d0[3] = {0, 0, 0}, d1[3] = {0, 0, 0};
      for (int i = 0; i < kK; ++i) for (int j = 0; j < 3; ++j) { d0[j] -= J0[6 * i + 3 + j]; d1[j] -= J1[6 * i + 3 + j]; }
      const M3 R = quat_matrix(Q4{body_q[4 * b], body_q[4 * b + 1], body_q[4 * b + 2], body_q[4 * b + 3]});
      const V3 v = v3(pw[3 * p] - body_t[3 * b], pw[3 * p + 1] - body_t[3 * b + 1], pw[3 * p + 2] - body_t[3 * b + 2]);
      double w0[9], w1[9];                               // world columns of the two residual rows: [rotation 3 | translation 3 | point 3]
      {
        const V3 c0 = cross(v, v3(d0[0], d0[1], d0[2])), c1 = cross(v, v3(d1[0], d1[1], d1[2]));
        w0[0] = 2.0 * c0.x; w0[1] = 2.0 * c0.y; w0[2] = 2.0 * c0.z; w1[0] = 2.0 * c1.x; w1[1] = 2.0 * c1.y; w1[2] = 2.0 * c1.z;
        for (int j = 0; j < 3; ++j) {
          w0[3 + j] = d0[j]; w1[3 + j] = d1[j];
          w0[6 + j] = d0[0] * R.m[j] + d0[1] * R.m[3 + j] + d0[2] * R.m[6 + j];
          w1[6 + j] = d1[0] * R.m[j] + d1[1] * R.m[3 + j] + d1[2] * R.m[6 + j];
        }
      }
      const int ds = seg - seg0;
      const bool in_smem = b == home_b && ds >= 0 && ds < kWorldMaxSeg;
      for (int a = 0; a < 9; ++a) {
        if (uw[a] < 0) continue;
        const bool sm = in_smem && a < 6;
        double* row = s_acc + a * kRow;                  // only used when sm
        auto add = [&](int slot, double* gptr, double val) { if (sm) atomicAdd(row + slot, val); else atomicAdd(gptr, val); };
        add(0, grad + n_a + uw[a], w0[a] * r0 + w1[a] * r1);
        for (int c = 0; c <= a; ++c) {                   // world x world (lower triangle within the block's own 9 columns)
          if (uw[c] < 0) continue;
          const double val = w0[a] * w0[c] + w1[a] * w1[c];
          if (sm && c < 6) atomicAdd(row + 1 + c, val);
          else {
            atomicAdd(Cmat + size_t(uw[a]) * N_c + uw[c], val);
            if (c != a) atomicAdd(Cmat + size_t(uw[c]) * N_c + uw[a], val);
          }
        }
        for (int j = 0; j < njc; ++j) {                  // world x the camera's calibration columns
          const double val = w0[a] * J0[kCpCols + j] + w1[a] * J1[kCpCols + j];
          const int uc = sd.calib_off + sd.junk[j];
          if (sm) atomicAdd(row + 7 + j, val);
          else { atomicAdd(Cmat + size_t(uw[a]) * N_c + uc, val); atomicAdd(Cmat + size_t(uc) * N_c + uw[a], val); }
        }
        for (int c = 0; c < kCpCols; ++c) {              // control points of the block's segment x world
          const double val = w0[a] * J0[c] + w1[a] * J1[c];
          if (sm) atomicAdd(row + 7 + kMaxCalib + ds * kCpCols + c, val);
          else atomicAdd(Bmat + (size_t(6) * seg + c) * N_c + uw[a], val);
        }
      }
    }
  }
  __syncthreads();
  // flush the tile's pose sums
  const int ur = body_u[2 * home_b], ut = body_u[2 * home_b + 1];
  for (int e = t; e < 6 * kRow; e += blockDim.x) {
    const double val = s_acc[e];
    if (val == 0.0) continue;
    const int a = e / kRow, k = e - a * kRow;
    const int ua = (a < 3 ? ur : ut) + a % 3;
    if (k == 0) atomicAdd(grad + n_a + ua, val);
    else if (k < 7) {
      const int c = k - 1, uc = (c < 3 ? ur : ut) + c % 3;
      atomicAdd(Cmat + size_t(ua) * N_c + uc, val);
      if (c != a) atomicAdd(Cmat + size_t(uc) * N_c + ua, val);
    } else if (k < 7 + kMaxCalib) {
      const int uc = sd.calib_off + sd.junk[k - 7];
      atomicAdd(Cmat + size_t(ua) * N_c + uc, val); atomicAdd(Cmat + size_t(uc) * N_c + ua, val);
    } else {
      const int q = k - 7 - kMaxCalib, ds = q / kCpCols, c = q - ds * kCpCols;
      atomicAdd(Bmat + (size_t(6) * (seg0 + ds) + c) * N_c + ua, val);
    }
  }
}

// Final control-point exchange: keep what this rank is responsible for, zero the rest, then sum across ranks.
__global__ void __launch_bounds__(256) mask_ctrl_kernel(long n_a, const unsigned char* __restrict__ cp_own, int count_shared,
                                                        const double* __restrict__ ctrl, double* __restrict__ out) {
  for (long i = long(blockIdx.x) * blockDim.x + threadIdx.x; i < n_a; i += long(gridDim.x) * blockDim.x) {
    const int own = cp_own[i / 6];
    out[i] = (own == kCpOwned || (own == kCpShared && count_shared)) ? ctrl[i] : 0.0;
  }
}

}  // namespace cb2
