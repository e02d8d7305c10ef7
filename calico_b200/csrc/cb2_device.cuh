// calico_b200 — device-side data layout shared by the host driver and the kernels.
//
// HBM layout (all FP64 unless noted; one allocation per array, resident for the lifetime of the uploaded problem):
//   ctrl[2][n_cp*6]            spline control points, current (x) and candidate point
//   state[2][n_sensors]        SensorState (intrinsics, extrinsics, latency, sigma, loss), current and candidate
//   knots[n_cp+6], basis[n_seg*36]   knot vector and per-segment 6x6 basis matrices (bspline.hpp:192-244)
//   pw[n_points*3]             world-frame model points R_wm p_m + t_wm (world model is constant: world_model.cpp:40-77)
//   frames[n_frames][44]       per camera image: pose-dependent quantities shared by its corners (FrameRec), rewritten by K0 every sweep
//   per sensor, observations SORTED BY SPLINE SEGMENT (cameras: then by stamp): stamp[n], meas[n*m], seg[n] (i32) / frm[n] (i32, camera), pt[n] (i32, camera),
//                              seg_start[n_seg+1] (CSR), r[n*m], J[n*m*jw]  (jw = 36 + enabled calibration columns)
//   normal equations           see cb2_normal.cu / cb2_schur.cu
#pragma once
#include "cb2_functors.cuh"

namespace cb2 {

// Observations per CTA in the residual/Jacobian sweep. Cameras: 128 (streaming, HBM-bound). IMU blocks are few (2 x 50 k at C4) and
// FP64-latency-bound: one warp per CTA keeps their shared-memory footprint small enough to slip between the camera CTAs of the
// concurrently running camera sweep instead of fencing whole SMs off.
CB2_HD constexpr int eval_tile(int kind) { return kind == 0 ? 128 : 32; }
// Field stride of the compact records in shared memory: odd, so that both the per-thread writes ([field][lane]) and the per-warp
// expansion reads ([lane -> field][obs]) are bank-conflict free.
CB2_HD constexpr int eval_rec_stride(int kind) { return eval_tile(kind) + 1; }
constexpr int kMaxCalib = 20;       // max calibration unknowns of one sensor: 12 intrinsics + 3 + 3 + 1
constexpr int kCpCols = 6 * kK;     // 36 control-point columns per residual block

struct SensorDesc {
  int kind, model, ni, m;
  int n_obs;                    // active (non-outlier) observations
  int calib_off, n_calib;       // slice of the calibration unknown vector owned by this sensor (tangent dims)
  int jw;                       // Jacobian columns stored per residual row: 36 + n_jcal
  int n_jcal;
  int jcanon[kMaxCalib];        // canonical column (see jac_entry) of stored calibration column j
  int junk[kMaxCalib];          // calibration-local unknown index of stored calibration column j
  int u_intr, u_rot, u_trans, u_lat;   // calibration-local unknown offsets, -1 when the block is constant
  const double* stamp; const double* meas; const int* seg; const int* pt;
  const int* seg_start;
  const int* frm;               // cameras: sensor-local image (frame) index of every observation, see camera_frame_kernel
  int frame_base;               // global index of this sensor's first image
  const int* frame_obs;         // cameras: [n_images + 1] first observation of every image (CSR over the sorted observations)
  const int* seg_frame;         // cameras: [n_seg + 1] first sensor-local image of every spline segment
  double* r; double* J; unsigned char* valid;
  // Cameras with n_calib <= 16: the Jacobian sweep itself forms, per image group of a CTA, the COMPACT Gram matrix of the image's residual
  // rows [g (6) | r | 0 | calibration (<= 16)] (24 x 24, six 8x8 DMMA tiles) and leaves
  //   per (CTA, image): slot gslot_base + image + CTA-in-sensor, kGramSlot doubles = [ [g | r] x g : 8 x 6 | calib 0..7 x g : 8 x 6 |
  //                     calib 8..15 x g : 8 x 6 | the image's basis weights w_0..w_5 | 1.0 (0.0 = unused slot) | pad ]  — everything
  //                     whose expansion needs the basis weights, self-contained so that the consumer has no further indirection;
  //   per (CTA, warp):  gcta[CTA-in-sensor][warp][kGramCta] = the calib x calib tiles (3 x 64) + the calibration gradient (16) summed over
  //                     the rows of the warp's image groups.
  // expand_gram_kernel / assemble_calib_kernel (cb2_normal.cuh) turn them into the normal equations, so the camera Jacobian is written
  // once and never read back. gslots == nullptr: this sensor goes through accumulate_kernel.
  double* gslots; int gslot_base;
  double* gcta;
  unsigned char gfield[2][24];  // record field of compact column c for residual row q (filled by the host with gram_field below)
};
constexpr int kGramSlot = 3 * 48 + 8;
constexpr int kGramSlotW = 3 * 48, kGramSlotFlag = 3 * 48 + 6;
constexpr int kGramCta = 3 * 64 + 16;

// ---- TMA (bulk asynchronous copy engine) 1-D global -> shared copies, completion tracked by an mbarrier in shared memory ----
// cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes (SASS UBLKCP): one thread issues the copy of a CONTIGUOUS block
// (16-byte aligned address and size), the data lands in shared memory without passing through registers, and every thread that
// needs it waits on the barrier's phase. The emulation build copies synchronously.
CB2_D void mbar_init(unsigned long long* bar, int arrivals) {
#if !defined(CB2_EMUL)
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(arrivals));
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
#else
  (void)bar; (void)arrivals;
#endif
}
CB2_D void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
#if !defined(CB2_EMUL)
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
#else
  (void)bar; (void)bytes;
#endif
}
CB2_D void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
#if !defined(CB2_EMUL)
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(dst_smem)), a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d), "l"(src_gmem), "r"(bytes), "r"(a) : "memory");
#else
  (void)bar;
  const unsigned char* s = static_cast<const unsigned char*>(src_gmem);
  unsigned char* d = static_cast<unsigned char*>(dst_smem);
  for (unsigned i = 0; i < bytes; ++i) d[i] = s[i];
#endif
}
CB2_D void mbar_wait(unsigned long long* bar, unsigned parity) {
#if !defined(CB2_EMUL)
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(a), "r"(parity) : "memory");
#else
  (void)bar; (void)parity;
#endif
}

// D(8x8) += A(8x4) B(4x8) in FP64 on the tensor pipe (mma.sync.m8n8k4.f64, SASS DMMA). Lane l holds a = A[l / 4][l % 4],
// b = B[l % 4][l / 4], c0, c1 = D[l / 4][2 (l % 4) + {0, 1}].
CB2_D void dmma_8x8x4(double& c0, double& c1, double a, double b) {
#if defined(CB2_EMUL)
  ::cb2emul::dmma_8x8x4(c0, c1, a, b);
#else
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
#endif
}



struct EvalTile { int sensor, start, count; };

// Scalars exchanged with the host once per LM iteration (device array of doubles). With several ranks, slots [0, 3) and
// [4, 11) are summed and slot 3 is max-reduced across ranks before the host reads them.
enum Scal {
  kScCost = 0, kScInvalid = 1,            // cost at x (1/2 sum rho), number of residual blocks that failed to evaluate
  kScGradSq = 2, kScGradMax = 3,          // |x - Plus(x, -g)|_2^2 and _inf  (trust_region_minimizer.cc, Ceres external)
  kScCandCost = 4, kScCandInvalid = 5,    // same as slots 0, 1 at the candidate point
  kScSolveFail = 6,                       // > 0 when a Cholesky pivot was not positive / step not finite
  kScModelChange = 7,                     // model cost change of the computed step
  kScStepNorm2 = 8, kScXNorm2 = 9, kScCandXNorm2 = 10,
  kScRadius = 11, kScLmLo = 12, kScLmHi = 13,   // host -> device: trust-region radius, min / max LM diagonal of the coming solve
  kScCount = 16
};

// Ownership of a control point on this rank (time-range sharding): interiors of owned chunks are updated and
// counted here; separators are replicated (updated everywhere, counted on rank 0 only); everything else belongs to a peer.
enum { kCpPeer = 0, kCpOwned = 1, kCpShared = 2 };

}  // namespace cb2
