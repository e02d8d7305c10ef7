// calico_b200 — sensor intrinsics models with analytic derivatives (host+device).
//
// Camera models restate calico/sensors/camera_models.h ProjectPoint bodies:
//   OpenCv5 :105-141, OpenCv8 :257-298, KannalaBrandt :420-462 (Taylor branch :444-446), DoubleSphere :623-657,
//   FieldOfView :740-781 (branches :762-772), UnifiedCamera :872-901, ExtendedUnifiedCamera :985-1015 (beta * norm, :995).
// IMU models restate accelerometer_models.h:80-85,129-141,208-235 / gyroscope_models.h:82,130,208.
// Enum values are ABI (camera_models.h:16-33, accelerometer_models.h:16-25).
#pragma once
#include "cb2_math.cuh"

namespace cb2 {

enum { kCamOpenCv5 = 1, kCamOpenCv8 = 2, kCamKannalaBrandt = 3, kCamDoubleSphere = 4, kCamFieldOfView = 5, kCamUnified = 6,
       kCamExtendedUnified = 7 };
enum { kImuScaleOnly = 1, kImuScaleAndBias = 2, kImuVectorNav = 3 };
constexpr int kMaxIntrinsics = 12;

CB2_HD int camera_num_params(int model) {
  switch (model) { case 1: return 8; case 2: return 11; case 3: return 7; case 4: return 5; case 5: return 4; case 6: return 4; case 7: return 5; default: return -1; }
}
CB2_HD int imu_num_params(int model) {
  switch (model) { case 1: return 1; case 2: return 4; case 3: return 12; default: return -1; }
}

// Projects p (camera frame) to a pixel. On success fills uv, d(uv)/dp as dp[2][3] and d(uv)/d(intrinsics) as di[2][ni].
// Returns false where the reference returns a non-OK status. If kJac is false the derivative outputs are left untouched.
template <bool kJac>
CB2_HD bool camera_project(int model, const double* in, const V3& p, double uv[2], double dp[2][3], double di[2][kMaxIntrinsics]) {
  const double f = in[0], cx = in[1], cy = in[2];
  switch (model) {
    case kCamOpenCv5:
    case kCamOpenCv8: {
      if (p.z <= 0.0) return false;
      const double iz = 1.0 / p.z;
      const double x = p.x * iz, y = p.y * iz;
      const double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
      const double k1 = in[3], k2 = in[4], p1 = in[5], p2 = in[6], k3 = in[7];
      const double num = 1.0 + r2 * (k1 + r2 * (k2 + r2 * k3));
      double s = num, den_inv = 1.0, den = 1.0;
      if (model == kCamOpenCv8) { den = 1.0 + r2 * (in[8] + r2 * (in[9] + r2 * in[10])); den_inv = 1.0 / den; s = num * den_inv; }
      const double px = x * s + 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x);
      const double py = y * s + 2.0 * p2 * x * y + p1 * (r2 + 2.0 * y * y);
      uv[0] = px * f + cx; uv[1] = py * f + cy;
      if (kJac) {
        double ds = k1 + r2 * (2.0 * k2 + 3.0 * r2 * k3);  // d num / d r2
        if (model == kCamOpenCv8) {
          const double dden = in[8] + r2 * (2.0 * in[9] + 3.0 * r2 * in[10]);
          ds = (ds - s * dden) * den_inv;
        }
        const double dpx_dx = s + 2.0 * x * x * ds + 2.0 * p1 * y + 6.0 * p2 * x;
        const double dpx_dy = 2.0 * x * y * ds + 2.0 * p1 * x + 2.0 * p2 * y;
        const double dpy_dx = 2.0 * x * y * ds + 2.0 * p2 * y + 2.0 * p1 * x;
        const double dpy_dy = s + 2.0 * y * y * ds + 2.0 * p2 * x + 6.0 * p1 * y;
        const double fz = f * iz;
        dp[0][0] = fz * dpx_dx; dp[0][1] = fz * dpx_dy; dp[0][2] = -fz * (dpx_dx * x + dpx_dy * y);
        dp[1][0] = fz * dpy_dx; dp[1][1] = fz * dpy_dy; dp[1][2] = -fz * (dpy_dx * x + dpy_dy * y);
        di[0][0] = px; di[1][0] = py;
        di[0][1] = 1.0; di[1][1] = 0.0;
        di[0][2] = 0.0; di[1][2] = 1.0;
        const double fx = f * x * den_inv, fy = f * y * den_inv;
        di[0][3] = fx * r2; di[1][3] = fy * r2;
        di[0][4] = fx * r4; di[1][4] = fy * r4;
        di[0][5] = f * 2.0 * x * y; di[1][5] = f * (r2 + 2.0 * y * y);
        di[0][6] = f * (r2 + 2.0 * x * x); di[1][6] = f * 2.0 * x * y;
        di[0][7] = fx * r6; di[1][7] = fy * r6;
        if (model == kCamOpenCv8) {
          const double gx = -f * x * s * den_inv, gy = -f * y * s * den_inv;
          di[0][8] = gx * r2; di[1][8] = gy * r2;
          di[0][9] = gx * r4; di[1][9] = gy * r4;
          di[0][10] = gx * r6; di[1][10] = gy * r6;
        }
      }
      return true;
    }
    case kCamKannalaBrandt: {
      if (p.z <= 0.0) return false;
      const double iz = 1.0 / p.z;
      const double x = p.x * iz, y = p.y * iz;
      const double r2 = x * x + y * y;
      const double r = sqrt(r2);
      const double k1 = in[3], k2 = in[4], k3 = in[5], k4 = in[6];
      double s, sx, sy;            // s and (ds/dx, ds/dy)
      double dk[4] = {0, 0, 0, 0}; // ds/dk_j
      if (r < 1e-9) {
        s = 1.0 + r2 * (k1 - 1.0 / 3.0 + r2 * (-k1 + k2 + 0.2));
        const double ds_dr2 = (k1 - 1.0 / 3.0) + 2.0 * r2 * (-k1 + k2 + 0.2);
        sx = 2.0 * x * ds_dr2; sy = 2.0 * y * ds_dr2;
        dk[0] = r2 - r2 * r2; dk[1] = r2 * r2;
      } else {
        const double th = atan(r), t2 = th * th;
        const double poly = 1.0 + t2 * (k1 + t2 * (k2 + t2 * (k3 + t2 * k4)));
        const double thd = th * poly;
        const double ir = 1.0 / r;
        s = thd * ir;
        const double dthd = 1.0 + t2 * (3.0 * k1 + t2 * (5.0 * k2 + t2 * (7.0 * k3 + t2 * 9.0 * k4)));
        const double ds_dr = (dthd / (1.0 + r2) - s) * ir;
        sx = ds_dr * x * ir; sy = ds_dr * y * ir;
        const double t3 = t2 * th * ir;
        dk[0] = t3; dk[1] = t3 * t2; dk[2] = t3 * t2 * t2; dk[3] = t3 * t2 * t2 * t2;
      }
      uv[0] = x * s * f + cx; uv[1] = y * s * f + cy;
      if (kJac) {
        const double fz = f * iz;
        const double a00 = s + x * sx, a01 = x * sy, a10 = y * sx, a11 = s + y * sy;
        dp[0][0] = fz * a00; dp[0][1] = fz * a01; dp[0][2] = -fz * (a00 * x + a01 * y);
        dp[1][0] = fz * a10; dp[1][1] = fz * a11; dp[1][2] = -fz * (a10 * x + a11 * y);
        di[0][0] = x * s; di[1][0] = y * s;
        di[0][1] = 1.0; di[1][1] = 0.0; di[0][2] = 0.0; di[1][2] = 1.0;
        for (int j = 0; j < 4; ++j) { di[0][3 + j] = f * x * dk[j]; di[1][3 + j] = f * y * dk[j]; }
      }
      return true;
    }
    case kCamDoubleSphere: {
      const double xi = in[3], al = in[4];
      const double w1 = al > 0.5 ? (1.0 - al) / al : al / (1.0 - al);
      const double num = w1 + xi;
      const double w2_sq = num * num / (2.0 * w1 * xi + xi * xi + 1.0);
      const double r2 = dot(p, p);
      if (p.z * p.z <= -w2_sq * r2) return false;
      const double r = sqrt(r2);
      const double d2 = r2 * (1.0 + xi * xi) + 2.0 * xi * r * p.z;
      const double d = sqrt(d2);
      const double den = al * d + (1.0 - al) * (xi * r + p.z);
      const double s = 1.0 / den;
      uv[0] = p.x * s * f + cx; uv[1] = p.y * s * f + cy;
      if (kJac) {
        // dr/dp = p/r ; dd/dp = ( (1+xi^2) p + xi (z p / r + r e_z) ) / d
        const double ir = 1.0 / r, id = 1.0 / d;
        double dden[3];
        for (int i = 0; i < 3; ++i) {
          const double pi = get(p, i);
          const double dr = pi * ir;
          const double dd = ((1.0 + xi * xi) * pi + xi * (p.z * dr + (i == 2 ? r : 0.0))) * id;
          dden[i] = al * dd + (1.0 - al) * (xi * dr + (i == 2 ? 1.0 : 0.0));
        }
        const double s2 = s * s;
        for (int i = 0; i < 3; ++i) {
          dp[0][i] = f * ((i == 0 ? s : 0.0) - p.x * s2 * dden[i]);
          dp[1][i] = f * ((i == 1 ? s : 0.0) - p.y * s2 * dden[i]);
        }
        di[0][0] = p.x * s; di[1][0] = p.y * s;
        di[0][1] = 1.0; di[1][1] = 0.0; di[0][2] = 0.0; di[1][2] = 1.0;
        const double dd_dxi = (xi * r2 + r * p.z) * id;
        const double dden_dxi = al * dd_dxi + (1.0 - al) * r;
        const double dden_dal = d - (xi * r + p.z);
        di[0][3] = -f * p.x * s2 * dden_dxi; di[1][3] = -f * p.y * s2 * dden_dxi;
        di[0][4] = -f * p.x * s2 * dden_dal; di[1][4] = -f * p.y * s2 * dden_dal;
      }
      return true;
    }
    case kCamFieldOfView: {
      const double w = in[3];
      if (p.z <= 0.0) return false;
      const double iz = 1.0 / p.z;
      const double x = p.x * iz, y = p.y * iz;
      const double r2 = x * x + y * y;
      const double r = sqrt(r2);
      double s, sx = 0.0, sy = 0.0, sw = 0.0;
      if (w * w < 1e-5) {
        s = 1.0;
      } else {
        const double th = tan(0.5 * w);
        const double tt = 2.0 * th;
        const double dtt = 1.0 + th * th;  // d(2 tan(w/2))/dw
        if (r2 < 1e-5) {
          s = tt / w;
          sw = (dtt - s) / w;
        } else {
          const double arg = r * tt;
          const double at = atan(arg);
          const double irw = 1.0 / (r * w);
          s = at * irw;
          const double dat = 1.0 / (1.0 + arg * arg);
          const double ds_dr = (dat * tt * irw) - s / r;
          sx = ds_dr * x / r; sy = ds_dr * y / r;
          sw = dat * r * dtt * irw - s / w;
        }
      }
      uv[0] = x * s * f + cx; uv[1] = y * s * f + cy;
      if (kJac) {
        const double fz = f * iz;
        const double a00 = s + x * sx, a01 = x * sy, a10 = y * sx, a11 = s + y * sy;
        dp[0][0] = fz * a00; dp[0][1] = fz * a01; dp[0][2] = -fz * (a00 * x + a01 * y);
        dp[1][0] = fz * a10; dp[1][1] = fz * a11; dp[1][2] = -fz * (a10 * x + a11 * y);
        di[0][0] = x * s; di[1][0] = y * s;
        di[0][1] = 1.0; di[1][1] = 0.0; di[0][2] = 0.0; di[1][2] = 1.0;
        di[0][3] = f * x * sw; di[1][3] = f * y * sw;
      }
      return true;
    }
    case kCamUnified:
    case kCamExtendedUnified: {
      const double al = in[3];
      const double w = al > 0.5 ? (1.0 - al) / al : al / (1.0 - al);
      double d, dd[3], dd_dbeta = 0.0;
      if (model == kCamUnified) {
        d = sqrt(dot(p, p));
        const double id = 1.0 / d;
        dd[0] = p.x * id; dd[1] = p.y * id; dd[2] = p.z * id;
      } else {
        const double beta = in[4];
        const double rho = sqrt(p.x * p.x + p.y * p.y);   // camera_models.h:995 — norm, not squared norm
        d = sqrt(beta * rho + p.z * p.z);
        const double id = 1.0 / d;
        const double irho = rho > 0.0 ? 1.0 / rho : 0.0;
        dd[0] = 0.5 * beta * p.x * irho * id; dd[1] = 0.5 * beta * p.y * irho * id; dd[2] = p.z * id;
        dd_dbeta = 0.5 * rho * id;
      }
      if (p.z <= -w * d) return false;
      const double den = al * d + (1.0 - al) * p.z;
      const double s = 1.0 / den;
      uv[0] = p.x * s * f + cx; uv[1] = p.y * s * f + cy;
      if (kJac) {
        const double s2 = s * s;
        for (int i = 0; i < 3; ++i) {
          const double dden = al * dd[i] + (i == 2 ? (1.0 - al) : 0.0);
          dp[0][i] = f * ((i == 0 ? s : 0.0) - p.x * s2 * dden);
          dp[1][i] = f * ((i == 1 ? s : 0.0) - p.y * s2 * dden);
        }
        di[0][0] = p.x * s; di[1][0] = p.y * s;
        di[0][1] = 1.0; di[1][1] = 0.0; di[0][2] = 0.0; di[1][2] = 1.0;
        const double dden_dal = d - p.z;
        di[0][3] = -f * p.x * s2 * dden_dal; di[1][3] = -f * p.y * s2 * dden_dal;
        if (model == kCamExtendedUnified) { di[0][4] = -f * p.x * s2 * al * dd_dbeta; di[1][4] = -f * p.y * s2 * al * dd_dbeta; }
      }
      return true;
    }
    default: return false;
  }
}

// IMU model: out = Project(in, w). dw = d out / d w (3x3), di = d out / d intrinsics (3 x ni).
template <bool kJac>
CB2_HD bool imu_project(int model, const double* in, const V3& w, V3* out, M3* dw, double di[3][kMaxIntrinsics]) {
  switch (model) {
    case kImuScaleOnly:
      *out = in[0] * w;
      if (kJac) { *dw = in[0] * m3_identity(); di[0][0] = w.x; di[1][0] = w.y; di[2][0] = w.z; }
      return true;
    case kImuScaleAndBias:
      *out = v3(in[0] * w.x + in[1], in[0] * w.y + in[2], in[0] * w.z + in[3]);
      if (kJac) {
        *dw = in[0] * m3_identity();
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) di[r][c] = 0.0;
        di[0][0] = w.x; di[1][0] = w.y; di[2][0] = w.z;
        di[0][1] = 1.0; di[1][2] = 1.0; di[2][3] = 1.0;
      }
      return true;
    case kImuVectorNav: {
      const double sx = in[0], sy = in[1], sz = in[2], a1 = in[3], a2 = in[4], a3 = in[5], a4 = in[6], a5 = in[7], a6 = in[8];
      const double ux = w.x + a1 * w.y + a2 * w.z, uy = w.y + a3 * w.x + a4 * w.z, uz = w.z + a5 * w.x + a6 * w.y;
      *out = v3(in[9] + sx * ux, in[10] + sy * uy, in[11] + sz * uz);
      if (kJac) {
        dw->m[0] = sx; dw->m[1] = sx * a1; dw->m[2] = sx * a2;
        dw->m[3] = sy * a3; dw->m[4] = sy; dw->m[5] = sy * a4;
        dw->m[6] = sz * a5; dw->m[7] = sz * a6; dw->m[8] = sz;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 12; ++c) di[r][c] = 0.0;
        di[0][0] = ux; di[1][1] = uy; di[2][2] = uz;
        di[0][3] = sx * w.y; di[0][4] = sx * w.z;
        di[1][5] = sy * w.x; di[1][6] = sy * w.z;
        di[2][7] = sz * w.x; di[2][8] = sz * w.y;
        di[0][9] = 1.0; di[1][10] = 1.0; di[2][11] = 1.0;
      }
      return true;
    }
    default: return false;
  }
}

}  // namespace cb2
