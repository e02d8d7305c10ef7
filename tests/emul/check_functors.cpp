// TEST HARNESS: compares the product's analytic Jacobians (calico_b200/csrc/cb2_functors.cuh, compiled for the host)
// with the oracle's dual-number Jacobians on random blocks. Build: see tests/test_functors_host.py.
#include <cstdio>
#include <random>
#include "../../oracle/calico_schur.hpp"
#include "cb2_functors.cuh"

using namespace cb2;

static double check_sensor(int kind, int model, unsigned seed, bool verbose, double theta_scale = 1.0) {
  std::mt19937_64 rng(seed);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  orc::Problem P;
  P.k = 6;
  const int n_cp = 14;
  const double dt = 0.1;
  for (int i = 0; i < n_cp + 6; ++i) P.knots.push_back(-5 * dt + dt * i);
  P.ctrl.resize(n_cp * 6);
  // smooth-ish control points around a base rotation of angle ~pi*theta_scale
  double base[3] = {0.3 * theta_scale, 2.9 * theta_scale, 0.4 * theta_scale};
  for (int i = 0; i < n_cp; ++i) {
    for (int d = 0; d < 3; ++d) P.ctrl[i * 6 + d] = base[d] + 0.2 * theta_scale * std::sin(0.7 * i + d) + 0.02 * theta_scale * U(rng);
    for (int d = 0; d < 3; ++d) P.ctrl[i * 6 + 3 + d] = (d == 2 ? 1.0 : 0.3) + 0.2 * std::cos(0.5 * i + d) + 0.02 * U(rng);
  }
  orc::RigidBody rb; rb.id = 0;
  const int npts = 9;
  for (int i = 0; i < npts; ++i) { rb.feature_ids.push_back(i); rb.pts.push_back(0.3 * (i % 3) + 0.1); rb.pts.push_back(0.3 * (i / 3) - 0.2); rb.pts.push_back(0.05 * U(rng)); }
  { double aa[3] = {0.1, -0.2, 0.15}; auto q = orc::AngleAxisToQuaternion(orc::V3<double>{aa[0], aa[1], aa[2]}); rb.q[0] = q.x; rb.q[1] = q.y; rb.q[2] = q.z; rb.q[3] = q.w; rb.t[0] = 0.05; rb.t[1] = -0.03; rb.t[2] = 0.02; }
  P.bodies.push_back(rb);
  orc::Sensor s; s.type = kind; s.model = model; s.name = "s";
  const int ni = kind == kCamera ? camera_num_params(model) : imu_num_params(model);
  s.intr.resize(ni);
  if (kind == kCamera) {
    s.intr[0] = 785; s.intr[1] = 640; s.intr[2] = 400;
    const double k5[5] = {-3.149e-1, 1.069e-1, 1.616e-4, 1.141e-4, -1.853e-2};
    if (model == 1 || model == 2) { for (int i = 0; i < 5; ++i) s.intr[3 + i] = k5[i]; if (model == 2) { s.intr[8] = 0.01; s.intr[9] = -0.02; s.intr[10] = 0.005; } }
    if (model == 3) { s.intr[3] = -3.149e-1; s.intr[4] = 1.069e-1; s.intr[5] = 1.616e-4; s.intr[6] = 1.141e-4; }
    if (model == 4) { s.intr[3] = -0.2; s.intr[4] = 0.55; }
    if (model == 5) { s.intr[3] = 0.9; }
    if (model == 6) { s.intr[3] = 0.6; }
    if (model == 7) { s.intr[3] = 0.6; s.intr[4] = 1.1; }
  } else {
    for (int i = 0; i < ni; ++i) s.intr[i] = 0.05 * U(rng);
    if (model == 1 || model == 2) s.intr[0] = 1.3;
    if (model == 3) { s.intr[0] = 1.1; s.intr[1] = 0.9; s.intr[2] = 1.2; }
  }
  { double aa[3] = {0.2 * U(rng), 0.2 * U(rng), 0.2 * U(rng)}; auto q = orc::AngleAxisToQuaternion(orc::V3<double>{aa[0], aa[1], aa[2]}); s.q[0] = q.x; s.q[1] = q.y; s.q[2] = q.z; s.q[3] = q.w; }
  for (int d = 0; d < 3; ++d) s.t[d] = 0.1 * U(rng);
  s.latency = 0.013; s.sigma = 0.37;
  s.en_intr = s.en_extr = s.en_lat = true;
  const int nobs = 40;
  for (int o = 0; o < nobs; ++o) {
    s.stamp.push_back(0.02 + 0.8 * (o + 0.5) / nobs);
    s.body_slot.push_back(0); s.feat_slot.push_back(o % npts);
    const int m = s.m();
    for (int q = 0; q < m; ++q) s.meas.push_back(kind == kCamera ? 500.0 + 100 * U(rng) : U(rng));
    s.outlier.push_back(0);
  }
  P.sensors.push_back(s);
  if (P.Build() != 0) { std::printf("build failed: %s\n", P.error.c_str()); return 1e9; }
  const orc::Sensor& S0 = P.sensors[0];
  SensorState st; st.kind = kind; st.model = model; st.ni = ni;
  for (int i = 0; i < ni; ++i) st.intr[i] = S0.intr[i];
  st.q = Q4{S0.q[0], S0.q[1], S0.q[2], S0.q[3]}; st.t = v3(S0.t[0], S0.t[1], S0.t[2]); st.latency = S0.latency; st.inv_sigma = 1.0 / S0.sigma;
  st.loss_type = 0; st.loss_scale = 1;
  const int m = S0.m();
  const int W = 36 + ni + 7;
  double worst = 0.0; int nvalid = 0;
  for (const auto& rbk : P.rblocks) {
    // oracle
    const int nb = int(rbk.blocks.size());
    std::vector<std::vector<double>> store(nb); double* jp[32]; double r_o[3];
    for (int i = 0; i < nb; ++i) { store[i].assign(size_t(m) * P.blocks[rbk.blocks[i]].tsize, 0.0); jp[i] = store[i].data(); }
    const bool ok_o = P.EvaluateBlock(rbk, r_o, jp);
    // product
    double recbuf[256]; Rec rec{recbuf, 1};
    const double* cp = &P.ctrl[size_t(rbk.seg.spline_index) * 6];
    bool ok_p;
    const int o = rbk.obs;
    if (kind == kCamera) {
      const orc::RigidBody& B = P.bodies[0];
      const orc::Qt<double> q{B.q[0], B.q[1], B.q[2], B.q[3]};
      const int f = S0.feat_slot[o];
      const orc::V3<double> pw = orc::q_rotate(q, orc::V3<double>{B.pts[3 * f], B.pts[3 * f + 1], B.pts[3 * f + 2]}) + orc::V3<double>{B.t[0], B.t[1], B.t[2]};
      ok_p = camera_block<true>(st, rbk.seg.basis, rbk.seg.knot0, rbk.seg.knot1, cp, rbk.seg.stamp, S0.meas[2 * o], S0.meas[2 * o + 1], v3(pw.x, pw.y, pw.z), rec);
    } else if (kind == kGyroscope) {
      ok_p = gyro_block<true>(st, rbk.seg.basis, rbk.seg.knot0, rbk.seg.knot1, cp, rbk.seg.stamp, v3(S0.meas[3 * o], S0.meas[3 * o + 1], S0.meas[3 * o + 2]), rec);
    } else {
      ok_p = accel_block<true>(st, v3(P.gravity[0], P.gravity[1], P.gravity[2]), rbk.seg.basis, rbk.seg.knot0, rbk.seg.knot1, cp, rbk.seg.stamp,
                               v3(S0.meas[3 * o], S0.meas[3 * o + 1], S0.meas[3 * o + 2]), rec);
    }
    if (ok_o != ok_p) { std::printf("validity mismatch obs %d: oracle %d product %d\n", o, ok_o, ok_p); worst = 1e9; continue; }
    if (!ok_o) continue;
    ++nvalid;
    const int cp_first = nb - 6;
    for (int q = 0; q < m; ++q) {
      double scale = 0.0;
      std::vector<double> Jo(W), Jp(W);
      for (int c = 0; c < 6; ++c) for (int d = 0; d < 6; ++d) Jo[c * 6 + d] = store[cp_first + c][q * 6 + d];
      for (int c = 0; c < ni; ++c) Jo[36 + c] = store[0][q * ni + c];
      for (int c = 0; c < 3; ++c) Jo[36 + ni + c] = store[1][q * 3 + c];
      for (int c = 0; c < 3; ++c) Jo[36 + ni + 3 + c] = store[2][q * 3 + c];
      Jo[36 + ni + 6] = store[3][q];
      for (int c = 0; c < W; ++c) { Jp[c] = jac_entry(kind, ni, rec, q, c); scale = std::max(scale, std::fabs(Jo[c])); }
      const double dr = std::fabs(rec.get(q) - r_o[q]) / std::max(1.0, std::fabs(r_o[q]));
      worst = std::max(worst, dr);
      for (int c = 0; c < W; ++c) {
        // relative to the magnitude of the column group the entry belongs to
        const double e = std::fabs(Jp[c] - Jo[c]) / std::max(1e-3 * scale, std::fabs(Jo[c]));
        if (e > 1e-7 && verbose) std::printf("  kind %d model %d obs %d row %d col %d: oracle % .12e product % .12e\n", kind, model, o, q, c, Jo[c], Jp[c]);
        worst = std::max(worst, e);
      }
    }
  }
  std::printf("kind %d model %d theta_scale %.3g: %d valid blocks, worst rel err %.3e\n", kind, model, theta_scale, nvalid, worst);
  return worst;
}

int main() {
  double worst = 0;
  for (int model = 1; model <= 7; ++model) worst = std::max(worst, check_sensor(kCamera, model, 11 + model, true));
  for (int model = 1; model <= 3; ++model) worst = std::max(worst, check_sensor(kGyroscope, model, 21 + model, true));
  for (int model = 1; model <= 3; ++model) worst = std::max(worst, check_sensor(kAccelerometer, model, 31 + model, true));
  // small rotation angles exercise the series branches
  for (double ts : {1e-2, 1e-4, 1e-9}) {
    worst = std::max(worst, check_sensor(kCamera, 1, 41, true, ts));
    worst = std::max(worst, check_sensor(kGyroscope, 2, 42, true, ts));
    worst = std::max(worst, check_sensor(kAccelerometer, 2, 43, true, ts));
  }
  std::printf("WORST %.3e\n", worst);
  return worst < 1e-7 ? 0 : 1;
}
