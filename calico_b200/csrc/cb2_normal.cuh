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
// Camera rows have more structure: all rows of one image share the basis weights w, and their control-point part is the Kronecker
// product g (x) w (g = d r / d pose, 6 numbers per row). Per image the kernel therefore multiplies only the COMPACT rows
// [g^ | r | calibration] (24 columns = 6 DMMA tiles instead of 28; g^ = g w_istar is read straight from the J columns of the control point
// with the largest weight) and expands once per image:  H[cp_a, cp_b] += (w_a w_b / w_istar^2) G^[.,.],  H[cp_a, r|calib] += (w_a / w_istar) ...
// This cuts the tensor work of K4 by ~4.7x, but measured on C4 it does not pay yet (378 us vs 361 us for the plain product): DRAM still
// fetches every sector of the J rows, the tile pipeline is one tile deep per warp, and the two warps that also own an IMU sensor become
// the critical path. It is therefore OPT-IN (CB2_ACC_STRUCTURED=1) until images are balanced across warps; parity-tested either way.
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
constexpr int kAccMaxSensors = 64;
constexpr int kAccMaxSlots = 16;   // camera images of one spline segment handled by the structured path (more: plain product for that segment)
constexpr int kAccSlot = 144;      // per image: G^gg 6x6 | sum r g^ (6) | sum c g^ (6 x 16) | w_a / w_istar (6)
constexpr int kAccSlotGr = 36, kAccSlotGc = 42, kAccSlotRho = 138;
CB2_HD constexpr size_t acc_smem_bytes(bool structured) {
  return (size_t(kAccWarps) * 2 * (kAccRows * kAccStride + 4) + (structured ? size_t(kAccMaxSlots) * kAccSlot : 0)) * sizeof(double);
}
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
// kStruct = false compiles the structured camera path out entirely (the default build of the hot kernel carries none of its cost).
template <int NB, bool kStruct>
__global__ void __launch_bounds__(kAccThreads, (NB == 7 ? CB2_ACC_MINBLOCKS : 2)) accumulate_kernel(
    const SensorDesc* __restrict__ sensors, int n_sensors, int N_c, int g_lo, const int* __restrict__ c2off, int csz, double* __restrict__ segA,
    double* __restrict__ segG, double* __restrict__ segB, double* __restrict__ segC, double* __restrict__ segGc, const double* __restrict__ frames) {
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
  bool dirty = false;
  __syncwarp();
  // Per-image slots of the structured camera accumulation (shared memory behind the tile buffers) and their per-sensor offsets.
  double* const slots = reinterpret_cast<double*>(tiles + kAccWarps);
  __shared__ int s_slot_begin[kAccMaxSensors + 1];
  __shared__ int s_use_struct;
  if (kStruct && warp == 0) {   // lane-parallel: one sensor per lane and round, exclusive prefix sum by shuffles (no serial chain of dependent loads)
    const bool ok = n_sensors <= kAccMaxSensors;
    int base = 0;
    for (int s0 = 0; ok && s0 < n_sensors; s0 += 32) {
      const int s = s0 + lane;
      int cnt = 0;
      if (s < n_sensors) {
        const SensorDesc& sd = sensors[s];
        if (sd.kind == kCamera && sd.seg_frame != nullptr && sd.n_calib <= 16) cnt = sd.seg_frame[g + 1] - sd.seg_frame[g];
      }
      int incl = cnt;
      for (int off = 1; off < 32; off <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += o; }
      if (s < n_sensors) s_slot_begin[s] = base + incl - cnt;
      base += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) { s_slot_begin[n_sensors <= kAccMaxSensors ? n_sensors : 0] = base; s_use_struct = (ok && base <= kAccMaxSlots) ? 1 : 0; }
  }
  if (kStruct) __syncthreads();
  const bool use_struct = kStruct && s_use_struct != 0;

  for (int s = warp; s < n_sensors; s += kAccWarps) {
    const SensorDesc& sd = sensors[s];
    const int m = sd.m, jw = sd.jw, nc = sd.n_calib;
    const int o0 = sd.seg_start[g];
    const int rows = (sd.seg_start[g + 1] - o0) * m;
    if (rows == 0) continue;
    if (kStruct && use_struct && sd.kind == kCamera && nc <= 16) {
      // ---- camera rows, image by image (see the header comment): compact Gram of [g^ | r | calibration]; the Kronecker expansion of its
      //      g^ rows is deferred to the end of the kernel (per-image slots in shared memory), the [r | calibration] square is summed here ----
      constexpr int CS = 28, CR = 16;                         // compact tile: 16 rows x 24 columns, row stride 28 (== 12 mod 16: conflict-free fragments)
      static_assert(CR * CS <= kAccRows * kAccStride + 4, "compact tile does not fit a tile buffer");
      double* const cb[2] = {tiles[warp][0], tiles[warp][1]};
      dirty = true;
      for (int i = lane; i < CR * CS; i += 32) { cb[0][i] = 0.0; cb[1][i] = 0.0; }
      __syncwarp();
      // lane -> the one compact column it copies: 0..5 g^ (J columns 6 istar + lane), 6 r, 8 + junk[j] calibration column j
      const int njc = jw - kCpCols;
      const int my_cal = lane - 8;
      const int dst_col = lane < 6 ? lane : (lane == 6 ? 6 : ((my_cal >= 0 && my_cal < njc) ? 8 + sd.junk[my_cal] : -1));
      const int f_begin = sd.seg_frame[g], f_end = sd.seg_frame[g + 1];
      // Per-image metadata is fetched once, up front (lane k: image f_begin + k), so that no dependent global load sits in the tile pipeline.
      const int nfr = f_end - f_begin;
      int my_obs0 = 0, my_rows = 0, my_istar = 0;
      if (lane < nfr) {
        my_obs0 = sd.frame_obs[f_begin + lane];
        my_rows = (sd.frame_obs[f_begin + lane + 1] - my_obs0) * 2;
        const double* w = frames + size_t(sd.frame_base + f_begin + lane) * FrameRec::kSize + FrameRec::w0;
        double bv = fabs(w[0]);
#pragma unroll
        for (int a = 1; a < kK; ++a) { const double v = fabs(w[a]); if (v > bv) { bv = v; my_istar = a; } }
      }
      auto item_rows = [&](int f) { return __shfl_sync(0xffffffffu, my_rows, (f - f_begin) & 31); };
      auto item_obs0 = [&](int f) { return __shfl_sync(0xffffffffu, my_obs0, (f - f_begin) & 31); };
      auto frame_istar = [&](int f) { return __shfl_sync(0xffffffffu, my_istar, (f - f_begin) & 31); };
      // flattened (image, pass) pipeline: the tile of the next item streams in while the current one is multiplied
      auto issue = [&](int f, int pass, int buf) {
        if (f < f_end) {
          const int nrows = item_rows(f), r0 = pass * CR, nr = min(CR, nrows - r0), istar = frame_istar(f);
          const size_t row0 = size_t(item_obs0(f)) * 2 + r0;                       // first J row of this tile
          double* tb = cb[buf];
          if (dst_col >= 0) {
            const double* src = lane == 6 ? sd.r + row0 : sd.J + row0 * jw + (lane < 6 ? 6 * istar + lane : kCpCols + my_cal);
            const int sstride = lane == 6 ? 1 : jw;
#pragma unroll 4
            for (int rr = 0; rr < nr; ++rr) cp_async8(tb + rr * CS + dst_col, src + size_t(rr) * sstride);
          }
          if (lane < 24) for (int rr = nr; rr < min(CR, (nr + 3) & ~3); ++rr) tb[rr * CS + lane] = 0.0;   // complete the last k-step with zero rows
        }
        cp_async_commit();
      };
      int cf = f_begin, cp = 0;                               // current item
      issue(cf, cp, 0);
      int buf = 0;
      double cg[6][2], cgS[6][2];                             // compact Gram of the current image / summed over the sensor's images
#pragma unroll
      for (int q = 0; q < 6; ++q) { cgS[q][0] = 0.0; cgS[q][1] = 0.0; }
      while (cf < f_end) {
        const int nrows = item_rows(cf), npass = (nrows + CR - 1) / CR;
        int nf = cf, np = cp + 1;
        if (np >= npass) { nf = cf + 1; np = 0; }
        issue(nf, np, buf ^ 1);
        cp_async_wait<1>();
        __syncwarp();
        if (cp == 0) {
#pragma unroll
          for (int q = 0; q < 6; ++q) { cg[q][0] = 0.0; cg[q][1] = 0.0; }
        }
        const double* tb = cb[buf];
        const int nr = min(CR, nrows - cp * CR);
#pragma unroll
        for (int ks = 0; ks < CR / 4; ++ks) {
          if (4 * ks < nr) {
            const double* row = tb + (4 * ks + fr) * CS + fc;
            const double f0 = row[0], f1 = row[8], f2 = row[16];
            dmma_8x8x4(cg[0][0], cg[0][1], f0, f0);
            dmma_8x8x4(cg[1][0], cg[1][1], f1, f0);
            dmma_8x8x4(cg[2][0], cg[2][1], f1, f1);
            dmma_8x8x4(cg[3][0], cg[3][1], f2, f0);
            dmma_8x8x4(cg[4][0], cg[4][1], f2, f1);
            dmma_8x8x4(cg[5][0], cg[5][1], f2, f2);
          }
        }
        __syncwarp();
        if (cp == npass - 1) {
          // image complete: its g^ rows go to the image's slot (expanded at the end of the kernel), the rest into the sensor sums
          double* slot = slots + size_t(s_slot_begin[s] + (cf - f_begin)) * kAccSlot;
          const double* w = frames + size_t(sd.frame_base + cf) * FrameRec::kSize + FrameRec::w0;
          const int istar = frame_istar(cf);
          if (lane < kK) slot[kAccSlotRho + lane] = w[lane] / w[istar];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int Jc = 2 * fr + i;                       // compact column of this accumulator entry (tiles (0,0), (1,0), (2,0))
            if (Jc < 6) {
              if (fc < 6) slot[fc * 6 + Jc] = cg[0][i];                           // G^gg[fc][Jc]
              else if (fc == 6) slot[kAccSlotGr + Jc] = cg[0][i];                 // sum r g^
              slot[kAccSlotGc + Jc * 16 + fc] = cg[1][i];                         // sum c_fc g^_Jc
              slot[kAccSlotGc + Jc * 16 + 8 + fc] = cg[3][i];
            }
          }
#pragma unroll
          for (int q = 0; q < 6; ++q) { cgS[q][0] += cg[q][0]; cgS[q][1] += cg[q][1]; }
        }
        cf = nf; cp = np; buf ^= 1;
      }
      cp_async_wait<0>();
      __syncwarp();
      // [r | calibration] x [r | calibration] of this sensor: calibration gradient and calibration block
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int Jc = 2 * fr + i;
        if (Jc == 6) {                                        // column r of tiles (1,0), (2,0)
          if (fc < nc) segGc[size_t(gl) * N_c + sd.calib_off + fc] = cgS[1][i];
          if (8 + fc < nc) segGc[size_t(gl) * N_c + sd.calib_off + 8 + fc] = cgS[3][i];
        }
        // tiles (1,1): (li, lj) = (fc, Jc); (2,1): (8 + fc, Jc); (2,2): (8 + fc, 8 + Jc)
        if (fc < nc && Jc <= fc) segC[size_t(gl) * csz + c2off[s] + fc * nc + Jc] = cgS[2][i];
        if (8 + fc < nc && Jc < nc) segC[size_t(gl) * csz + c2off[s] + (8 + fc) * nc + Jc] = cgS[4][i];
        if (8 + fc < nc && Jc <= fc) segC[size_t(gl) * csz + c2off[s] + (8 + fc) * nc + 8 + Jc] = cgS[5][i];
      }
      continue;
    } else {
    if (dirty) {   // a camera pass used these buffers with its own layout: restore the all-zero padding the generic path relies on
      for (int i = lane; i < 2 * (kAccRows * kAccStride + 4); i += 32) tb0[i] = 0.0;
      dirty = false;
      __syncwarp();
    }
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
    }
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
    if (kStruct && use_struct) {   // Kronecker expansion of the camera images: (w_a w_b / w_istar^2) G^gg[p][q], (w_b / w_istar) sum r g^_q
      const int nslots = s_slot_begin[n_sensors];
      const int a6 = I / 6, p6 = I - 6 * a6, b6 = Jx / 6, q6 = Jx - 6 * b6;
      for (int k = 0; k < nslots; ++k) {
        const double* slot = slots + size_t(k) * kAccSlot;
        if (I < kCpCols) v += slot[kAccSlotRho + a6] * slot[kAccSlotRho + b6] * slot[p6 * 6 + q6];
        else v += slot[kAccSlotRho + b6] * slot[kAccSlotGr + q6];
      }
    }
    if (I < kCpCols) segA[(size_t(gl) * kCpCols + I) * kCpCols + Jx] = v;
    else segG[size_t(gl) * kCpCols + Jx] = v;
  }
  if (kStruct && use_struct) {     // control points x calibration of every structured camera: (w_a / w_istar) sum c g^_p over its images
    for (int s = 0; s < n_sensors; ++s) {
      const int k0 = s_slot_begin[s], k1 = s_slot_begin[s + 1];
      if (k1 == k0) continue;
      const SensorDesc& sd = sensors[s];
      const int nc = sd.n_calib;
      for (int e = t; e < kCpCols * 16; e += kAccThreads) {
        const int Jx = e >> 4, li = e & 15, a6 = Jx / 6, p6 = Jx - 6 * a6;
        if (li >= nc) continue;
        double v = 0.0;
        for (int k = k0; k < k1; ++k) { const double* slot = slots + size_t(k) * kAccSlot; v += slot[kAccSlotRho + a6] * slot[kAccSlotGc + p6 * 16 + li]; }
        segB[(size_t(gl) * kCpCols + Jx) * N_c + sd.calib_off + li] = v;
      }
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
