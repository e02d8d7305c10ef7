// ORACLE — TEST INFRASTRUCTURE ONLY (see jet.hpp header).
//
// Restatement of the Ceres trust-region minimizer with the Levenberg-Marquardt strategy, as
// selected by calico::DefaultSolverOptions() + ceres::Solve (batch_optimizer.cpp:10-17,73).
// Ceres is external to the reference tree; function names below refer to Ceres 2.1/2.2's
// internal/ceres/{trust_region_minimizer,levenberg_marquardt_strategy,trust_region_step_evaluator,
// corrector,program_evaluator}.cc. See calico_problem.hpp for the parity status of this file.
#pragma once
#include "calico_problem.hpp"

namespace orc {

inline double NowSeconds() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Dense Cholesky (lower, in place, row-major n x n). Returns false if not positive definite.
// Right-looking, blocked (panels of kCholNB columns), OpenMP over the tiles of the trailing SYRK update, register-tiled micro-kernel on
// GCC vector extensions (8 rows x 24 columns of accumulators; AVX-512 or 2 x AVX2 with -march=native). This is the CPU baseline's dominant
// kernel on the BASELINE shapes (Ceres's DENSE_SCHUR factors a dense reduced system of ~5 n_cp unknowns with Eigen's / LAPACK's blocked LLT),
// so it is written to run near the host's FP64 peak rather than as a textbook triple loop: the reported CPU baseline is not a straw man.
typedef double CholVec __attribute__((vector_size(64), aligned(8)));
constexpr int kCholNB = 128, kCholMR = 8, kCholNRV = 3, kCholNR = 8 * kCholNRV;

// C[i0 .. i0+MR)[j0 .. j0+NR) -= sum_k A[i][kb + k] * Bt[k][j - jbase]; only entries with j <= i are stored (lower triangle).
inline void CholMicroKernel(double* A, int n, int kb, int nk, const double* Bt, int ldb, int jbase, int i0, int j0, int mr, int nr) {
  CholVec acc[kCholMR][kCholNRV];
  for (int r = 0; r < kCholMR; ++r) for (int v = 0; v < kCholNRV; ++v) acc[r][v] = CholVec{0, 0, 0, 0, 0, 0, 0, 0};
  const double* a[kCholMR];
  for (int r = 0; r < kCholMR; ++r) a[r] = A + size_t(i0 + std::min(r, mr - 1)) * n + kb;
  const double* bt = Bt + (j0 - jbase);
  for (int k = 0; k < nk; ++k) {
    CholVec b[kCholNRV];
    for (int v = 0; v < kCholNRV; ++v) b[v] = *reinterpret_cast<const CholVec*>(bt + size_t(k) * ldb + 8 * v);
#pragma GCC unroll 8
    for (int r = 0; r < kCholMR; ++r) {
      const double ar = a[r][k];
      const CholVec av = {ar, ar, ar, ar, ar, ar, ar, ar};
      for (int v = 0; v < kCholNRV; ++v) acc[r][v] += av * b[v];
    }
  }
  for (int r = 0; r < mr; ++r) {
    const int i = i0 + r;
    double* c = A + size_t(i) * n + j0;
    const int jn = std::min(nr, i - j0 + 1);
    for (int jj = 0; jj < jn; ++jj) c[jj] -= acc[r][jj >> 3][jj & 7];
  }
}

inline bool DenseCholesky(double* A, int n) {
  std::vector<double> Bt;
  for (int kb = 0; kb < n; kb += kCholNB) {
    const int ke = std::min(n, kb + kCholNB), nk = ke - kb;
    // Diagonal block, unblocked.
    for (int j = kb; j < ke; ++j) {
      double* aj = A + size_t(j) * n;
      double d = aj[j];
      for (int t = kb; t < j; ++t) d -= aj[t] * aj[t];
      if (!(d > 0.0) || !std::isfinite(d)) return false;
      d = std::sqrt(d);
      aj[j] = d;
      const double inv = 1.0 / d;
      for (int i = j + 1; i < ke; ++i) {
        double* ai = A + size_t(i) * n;
        double s = ai[j];
        for (int t = kb; t < j; ++t) s -= ai[t] * aj[t];
        ai[j] = s * inv;
      }
    }
    const int nrem = n - ke;
    if (nrem <= 0) break;
    // Panel: rows below the diagonal block, L21 = A21 L11^-T (one forward substitution per row), and its transposed, padded copy Bt.
    const int ldb = (nrem + kCholNR + 7) / 8 * 8;
    Bt.assign(size_t(nk) * ldb, 0.0);
#pragma omp parallel for schedule(static)
    for (int i = ke; i < n; ++i) {
      double* ai = A + size_t(i) * n;
      for (int j = kb; j < ke; ++j) {
        const double* aj = A + size_t(j) * n;
        double s = ai[j];
        for (int t = kb; t < j; ++t) s -= ai[t] * aj[t];
        ai[j] = s / aj[j];
      }
      for (int k = 0; k < nk; ++k) Bt[size_t(k) * ldb + (i - ke)] = ai[kb + k];
    }
    // Trailing update of the lower triangle: tiles of (64 rows) x (all columns up to the diagonal), MR x NR register tiles inside.
    const int IB = 64;
    const int nib = (nrem + IB - 1) / IB;
#pragma omp parallel for schedule(dynamic, 1)
    for (int ibr = 0; ibr < nib; ++ibr) {
      const int ib = nib - 1 - ibr;                       // longest rows first
      const int ia = ke + ib * IB, iz = std::min(n, ia + IB);
      for (int j0 = ke; j0 < iz; j0 += kCholNR) {
        const int nr = std::min(kCholNR, iz - j0);
        for (int i0 = std::max(ia, j0 / 1); i0 < iz; i0 += kCholMR) {
          if (i0 + kCholMR - 1 < j0) continue;            // tile entirely above the diagonal
          CholMicroKernel(A, n, kb, nk, Bt.data(), ldb, ke, i0, j0, std::min(kCholMR, iz - i0), nr);
        }
      }
    }
  }
  return true;
}
// Solves L L^T x = b in place; both sweeps walk the factor row by row (contiguous).
inline void CholeskySolve(const double* L, int n, double* b) {
  for (int i = 0; i < n; ++i) { double s = b[i]; const double* li = L + size_t(i) * n; for (int t = 0; t < i; ++t) s -= li[t] * b[t]; b[i] = s / li[i]; }
  for (int i = n - 1; i >= 0; --i) {
    const double* li = L + size_t(i) * n;
    const double x = b[i] / li[i];
    b[i] = x;
    for (int t = 0; t < i; ++t) b[t] -= li[t] * x;
  }
}

struct Minimizer {
  Problem& p;
  Options opt;
  Summary* summary;
  std::vector<IterationLog>* log;

  // Reduced program.
  std::vector<int> active_blocks;       // block ids in the reduced program
  std::vector<int> active_rblocks;      // residual block ids in the reduced program
  int n_tan = 0, n_amb = 0, n_res = 0;
  int n_cp_tan = 0;                      // control-point unknowns come first in the tangent ordering

  // Block-sparse Jacobian: per active residual block, per active parameter block an m x t tile.
  struct Tile { int block; int off; int t; size_t val; };
  std::vector<std::vector<Tile>> tiles;  // per active residual block
  std::vector<int> row0;                 // first scalar row of each active residual block
  std::vector<double> jvals, residuals, gradient, scaling;

  explicit Minimizer(Problem& prob, const Options& o, Summary* s, std::vector<IterationLog>* l)
      : p(prob), opt(o), summary(s), log(l) {}

  // Program::RemoveFixedBlocks (reduced_program.cc): constant or unreferenced parameter blocks drop
  // out; residual blocks depending only on such blocks are evaluated once into fixed_cost.
  bool Setup() {
    for (auto& b : p.blocks) { b.referenced = false; b.off = b.aoff = -1; }
    for (const auto& rb : p.rblocks) for (int id : rb.blocks) p.blocks[id].referenced = true;
    n_tan = n_amb = 0;
    active_blocks.clear();
    // Tangent ordering: control points first (they are the Schur e-blocks), then the rest.
    for (int pass = 0; pass < 2; ++pass)
      for (size_t i = 0; i < p.blocks.size(); ++i) {
        ParamBlock& b = p.blocks[i];
        if (b.constant || !b.referenced) continue;
        if ((pass == 0) != b.is_control_point) continue;
        b.off = n_tan; b.aoff = n_amb; n_tan += b.tsize; n_amb += b.size;
        active_blocks.push_back(int(i));
        if (pass == 0) n_cp_tan = n_tan;
      }
    summary->num_parameter_blocks = int(p.blocks.size());
    summary->num_parameters = summary->num_effective_parameters = 0;
    for (const auto& b : p.blocks) { summary->num_parameters += b.size; summary->num_effective_parameters += b.tsize; }
    summary->num_residual_blocks = int(p.rblocks.size());
    summary->num_residuals = 0;
    for (const auto& rb : p.rblocks) summary->num_residuals += rb.m;
    summary->num_parameter_blocks_reduced = int(active_blocks.size());
    summary->num_parameters_reduced = n_amb;
    summary->num_effective_parameters_reduced = n_tan;
    active_rblocks.clear(); tiles.clear(); row0.clear();
    n_res = 0;
    summary->fixed_cost = 0.0;
    size_t nvals = 0;
    for (size_t i = 0; i < p.rblocks.size(); ++i) {
      const ResidualBlock& rb = p.rblocks[i];
      std::vector<Tile> tl;
      for (int id : rb.blocks) {
        const ParamBlock& b = p.blocks[id];
        if (b.off < 0) continue;
        bool dup = false; for (auto& t : tl) dup |= (t.block == id);
        if (dup) continue;
        tl.push_back({id, b.off, b.tsize, nvals});
        nvals += size_t(rb.m) * b.tsize;
      }
      if (tl.empty()) {
        double r[3];
        if (!p.EvaluateBlock(rb, r, nullptr)) return false;
        double sq = 0; for (int q = 0; q < rb.m; ++q) sq += r[q] * r[q];
        double rho[3]; const Sensor& s = p.sensors[rb.sensor];
        EvaluateLoss(s.loss_type, s.loss_scale, sq, rho);
        summary->fixed_cost += 0.5 * rho[0];
        continue;
      }
      active_rblocks.push_back(int(i));
      tiles.push_back(std::move(tl));
      row0.push_back(n_res);
      n_res += rb.m;
    }
    summary->num_residual_blocks_reduced = int(active_rblocks.size());
    summary->num_residuals_reduced = n_res;
    jvals.assign(nvals, 0.0);
    residuals.assign(n_res, 0.0);
    gradient.assign(n_tan, 0.0);
    return true;
  }

  // ProgramEvaluator::Evaluate + ResidualBlock::Evaluate + Corrector (Ceres external).
  bool Evaluate(double* cost, bool with_jacobian) {
    double total = 0.0;
    bool ok = true;
    const int nrb = int(active_rblocks.size());
#pragma omp parallel for schedule(static) reduction(+ : total) num_threads(opt.num_threads)
    for (int ai = 0; ai < nrb; ++ai) {
      if (!ok) continue;
      const ResidualBlock& rb = p.rblocks[active_rblocks[ai]];
      const Sensor& s = p.sensors[rb.sensor];
      double r[3];
      double* jac[32];
      if (with_jacobian) {
        for (size_t i = 0; i < rb.blocks.size(); ++i) {
          jac[i] = nullptr;
          for (const auto& t : tiles[ai]) if (t.block == rb.blocks[i]) jac[i] = jvals.data() + t.val;
          // A block listed twice receives its Jacobian only once (cannot happen for these functors).
          for (size_t j = 0; j < i; ++j) if (rb.blocks[j] == rb.blocks[i]) jac[i] = nullptr;
        }
      }
      if (!p.EvaluateBlock(rb, r, with_jacobian ? jac : nullptr)) {
#pragma omp atomic write
        ok = false;
        continue;
      }
      double sq = 0.0; for (int q = 0; q < rb.m; ++q) sq += r[q] * r[q];
      double rho[3];
      EvaluateLoss(s.loss_type, s.loss_scale, sq, rho);
      total += 0.5 * rho[0];
      if (with_jacobian) {
        if (s.loss_type != kLossNone) {
          // Corrector::Corrector (corrector.cc): sq_norm == 0 or rho'' <= 0 → plain sqrt(rho') scaling.
          const double sqrt_rho1 = std::sqrt(rho[1]);
          double residual_scaling, alpha_sq_norm;
          if (sq == 0.0 || rho[2] <= 0.0) { residual_scaling = sqrt_rho1; alpha_sq_norm = 0.0; }
          else {
            const double D = 1.0 + 2.0 * sq * rho[2] / rho[1];
            const double alpha = 1.0 - std::sqrt(D);
            residual_scaling = sqrt_rho1 / (1 - alpha);
            alpha_sq_norm = alpha / sq;
          }
          for (const auto& t : tiles[ai]) {
            double* J = jvals.data() + t.val;
            if (alpha_sq_norm == 0.0) { for (int e = 0; e < rb.m * t.t; ++e) J[e] *= sqrt_rho1; }
            else {
              for (int c = 0; c < t.t; ++c) {
                double rtj = 0.0; for (int q = 0; q < rb.m; ++q) rtj += J[q * t.t + c] * r[q];
                for (int q = 0; q < rb.m; ++q) J[q * t.t + c] = sqrt_rho1 * (J[q * t.t + c] - alpha_sq_norm * r[q] * rtj);
              }
            }
          }
          for (int q = 0; q < rb.m; ++q) r[q] *= residual_scaling;
        }
        for (int q = 0; q < rb.m; ++q) residuals[row0[ai] + q] = r[q];
      }
    }
    if (!ok) return false;
    *cost = total;
    if (with_jacobian) {
      std::fill(gradient.begin(), gradient.end(), 0.0);
      for (int ai = 0; ai < nrb; ++ai) {
        const int m = p.rblocks[active_rblocks[ai]].m;
        for (const auto& t : tiles[ai]) {
          const double* J = jvals.data() + t.val;
          for (int q = 0; q < m; ++q) for (int c = 0; c < t.t; ++c) gradient[t.off + c] += J[q * t.t + c] * residuals[row0[ai] + q];
        }
      }
    }
    return true;
  }

  void SquaredColumnNorm(std::vector<double>& out) const {
    out.assign(n_tan, 0.0);
    for (size_t ai = 0; ai < tiles.size(); ++ai) {
      const int m = p.rblocks[active_rblocks[ai]].m;
      for (const auto& t : tiles[ai]) {
        const double* J = jvals.data() + t.val;
        for (int q = 0; q < m; ++q) for (int c = 0; c < t.t; ++c) out[t.off + c] += J[q * t.t + c] * J[q * t.t + c];
      }
    }
  }
  void ScaleColumns(const std::vector<double>& s) {
    for (size_t ai = 0; ai < tiles.size(); ++ai) {
      const int m = p.rblocks[active_rblocks[ai]].m;
      for (const auto& t : tiles[ai]) {
        double* J = jvals.data() + t.val;
        for (int q = 0; q < m; ++q) for (int c = 0; c < t.t; ++c) J[q * t.t + c] *= s[t.off + c];
      }
    }
  }

  // State vector helpers (reduced ambient vector).
  void GetState(std::vector<double>& x) const {
    x.resize(n_amb);
    for (int id : active_blocks) { const ParamBlock& b = p.blocks[id]; std::memcpy(&x[b.aoff], b.ptr, sizeof(double) * b.size); }
  }
  void SetState(const std::vector<double>& x) {
    for (int id : active_blocks) { const ParamBlock& b = p.blocks[id]; std::memcpy(b.ptr, &x[b.aoff], sizeof(double) * b.size); }
  }
  void Plus(const std::vector<double>& x, const std::vector<double>& delta, std::vector<double>& out) const {
    out.resize(n_amb);
    for (int id : active_blocks) {
      const ParamBlock& b = p.blocks[id];
      if (b.quaternion) QuaternionPlus(&x[b.aoff], &delta[b.off], &out[b.aoff]);
      else for (int i = 0; i < b.size; ++i) out[b.aoff + i] = x[b.aoff + i] + delta[b.off + i];
    }
  }

  // ---- linear solvers: minimise |J y - r|^2 + |D y|^2, i.e. (J^T J + D^2) y = J^T r. ----
  bool SolveDenseNormal(const std::vector<double>& D, std::vector<double>& y) {
    const int n = n_tan;
    std::vector<double> H(size_t(n) * n, 0.0);
    for (size_t ai = 0; ai < tiles.size(); ++ai) {
      const int m = p.rblocks[active_rblocks[ai]].m;
      const auto& tl = tiles[ai];
      for (size_t a = 0; a < tl.size(); ++a) for (size_t b = 0; b < tl.size(); ++b) {
        if (tl[b].off > tl[a].off) continue;  // lower triangle only (row block a >= col block b)
        const double* Ja = jvals.data() + tl[a].val; const double* Jb = jvals.data() + tl[b].val;
        for (int i = 0; i < tl[a].t; ++i) for (int j = 0; j < tl[b].t; ++j) {
          double s = 0.0; for (int q = 0; q < m; ++q) s += Ja[q * tl[a].t + i] * Jb[q * tl[b].t + j];
          H[size_t(tl[a].off + i) * n + tl[b].off + j] += s;
        }
      }
    }
    y.assign(n, 0.0);
    for (size_t ai = 0; ai < tiles.size(); ++ai) {
      const int m = p.rblocks[active_rblocks[ai]].m;
      for (const auto& t : tiles[ai]) { const double* J = jvals.data() + t.val; for (int q = 0; q < m; ++q) for (int c = 0; c < t.t; ++c) y[t.off + c] += J[q * t.t + c] * residuals[row0[ai] + q]; }
    }
    for (int i = 0; i < n; ++i) H[size_t(i) * n + i] += D[i] * D[i];
    if (!DenseCholesky(H.data(), n)) return false;
    CholeskySolve(H.data(), n, y.data());
    return true;
  }

  bool LinearSolve(const std::vector<double>& D, std::vector<double>& y);  // dispatch, defined in calico_schur.hpp

  static void Format(char* dst, size_t n, const char* fmt, double a = 0, double b = 0) { std::snprintf(dst, n, fmt, a, b); }

  // TrustRegionMinimizer::Minimize.
  void Minimize() {
    const double t_start = NowSeconds();
    Summary& S = *summary;
    if (!Setup()) { S.termination_type = kFailure; Format(S.message, sizeof S.message, "Initial residual and Jacobian evaluation failed."); return; }
    if (n_tan == 0) {
      // ceres::Solve on a program with no free parameters: "Function tolerance reached. No non-constant parameter blocks found."
      double c = 0; Evaluate(&c, false);
      S.initial_cost = S.final_cost = S.fixed_cost; S.termination_type = kConvergence;
      Format(S.message, sizeof S.message, "Function tolerance reached. No non-constant parameter blocks found.");
      return;
    }
    std::vector<double> x, candidate_x, delta(n_tan), step(n_tan), neg_grad(n_tan), proj, diag, lm_diag(n_tan);
    GetState(x);
    double x_norm = 0; for (double v : x) x_norm += v * v; x_norm = std::sqrt(x_norm);
    double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
    bool reuse_diagonal = false;
    double x_cost = 0, candidate_cost = 0, model_cost_change = 0;
    int num_consecutive_invalid_steps = 0;
    scaling.assign(n_tan, 1.0);
    IterationLog it{};

    auto evaluate_gradient_and_jacobian = [&](bool first) -> bool {
      const double t0 = NowSeconds();
      if (!Evaluate(&x_cost, true)) return false;
      it.cost = x_cost + S.fixed_cost;
      if (opt.jacobi_scaling) {
        if (first) { SquaredColumnNorm(scaling); for (auto& v : scaling) v = 1.0 / (1.0 + std::sqrt(v)); }
        ScaleColumns(scaling);
      }
      for (int i = 0; i < n_tan; ++i) neg_grad[i] = -gradient[i];
      Plus(x, neg_grad, proj);
      double mx = 0, nn = 0;
      for (int i = 0; i < n_amb; ++i) { const double d = x[i] - proj[i]; mx = std::max(mx, std::fabs(d)); nn += d * d; }
      it.gradient_max_norm = mx; it.gradient_norm = std::sqrt(nn);
      S.jacobian_time += NowSeconds() - t0;
      return true;
    };

    // IterationZero.
    double t_iter = NowSeconds();
    it.iteration = 0; it.trust_region_radius = radius;
    if (!evaluate_gradient_and_jacobian(true)) {
      S.termination_type = kFailure; Format(S.message, sizeof S.message, "Initial residual and Jacobian evaluation failed.");
      return;
    }
    S.initial_cost = x_cost + S.fixed_cost;
    it.step_is_valid = 1; it.step_is_successful = 1;
    // TrustRegionStepEvaluator with max_consecutive_nonmonotonic_steps = 0 reduces to the plain ratio.
    double reference_cost = x_cost;
    std::vector<IterationLog> local_log;
    std::vector<IterationLog>& L = log ? *log : local_log;
    L.clear();
    if (opt.minimizer_progress_to_stdout)
      std::printf("iter      cost      cost_change  |gradient|   |step|    tr_ratio  tr_radius  ls_iter  iter_time  total_time\n");

    for (;;) {
      // FinalizeIterationAndCheckIfMinimizerCanContinue.
      it.trust_region_radius = radius;
      it.iteration_time = NowSeconds() - t_iter;
      L.push_back(it);
      if (opt.minimizer_progress_to_stdout)
        std::printf("% 4d % 8e   % 3.2e   % 3.2e  % 3.2e  % 3.2e % 3.2e     % 4d   % 3.2e   % 3.2e\n", it.iteration, it.cost, it.cost_change,
                    it.gradient_max_norm, it.step_norm, it.relative_decrease, it.trust_region_radius, 1, it.iteration_time, NowSeconds() - t_start);
      if (it.iteration >= opt.max_num_iterations) { S.termination_type = kNoConvergence; Format(S.message, sizeof S.message, "Maximum number of iterations reached. Number of iterations: %.0f.", it.iteration); break; }
      if (it.step_is_successful && it.gradient_max_norm <= opt.gradient_tolerance) {
        S.termination_type = kConvergence; Format(S.message, sizeof S.message, "Gradient tolerance reached. Gradient max norm: %e <= %e", it.gradient_max_norm, opt.gradient_tolerance); break;
      }
      if (radius <= opt.min_trust_region_radius) {
        S.termination_type = kConvergence; Format(S.message, sizeof S.message, "Minimum trust region radius reached. Trust region radius: %e <= %e", radius, opt.min_trust_region_radius); break;
      }
      t_iter = NowSeconds();
      const double prev_gmax = it.gradient_max_norm, prev_gnorm = it.gradient_norm;
      const int next_iter = it.iteration + 1;
      it = IterationLog{};
      it.iteration = next_iter; it.gradient_max_norm = prev_gmax; it.gradient_norm = prev_gnorm;

      // ComputeTrustRegionStep → LevenbergMarquardtStrategy::ComputeStep.
      const double t_ls = NowSeconds();
      if (!reuse_diagonal) {
        SquaredColumnNorm(diag);
        for (auto& v : diag) v = std::min(std::max(v, opt.min_lm_diagonal), opt.max_lm_diagonal);
      }
      for (int i = 0; i < n_tan; ++i) lm_diag[i] = std::sqrt(diag[i] / radius);
      bool solved = LinearSolve(lm_diag, step);
      if (solved) for (int i = 0; i < n_tan; ++i) if (!std::isfinite(step[i])) solved = false;
      reuse_diagonal = true;
      S.linear_solver_time += NowSeconds() - t_ls;
      it.step_is_valid = 0;
      if (solved) {
        for (auto& v : step) v = -v;
        // model_cost_change = -(J step)^T (r + J step / 2).
        std::vector<double> model(n_res, 0.0);
        for (size_t ai = 0; ai < tiles.size(); ++ai) {
          const int m = p.rblocks[active_rblocks[ai]].m;
          for (const auto& t : tiles[ai]) { const double* J = jvals.data() + t.val; for (int q = 0; q < m; ++q) { double s = 0; for (int c = 0; c < t.t; ++c) s += J[q * t.t + c] * step[t.off + c]; model[row0[ai] + q] += s; } }
        }
        model_cost_change = 0.0;
        for (int i = 0; i < n_res; ++i) model_cost_change -= model[i] * (residuals[i] + model[i] / 2.0);
        it.step_is_valid = model_cost_change > 0.0;
        if (it.step_is_valid) { for (int i = 0; i < n_tan; ++i) delta[i] = step[i] * scaling[i]; num_consecutive_invalid_steps = 0; }
      }
      if (!it.step_is_valid) {
        // HandleInvalidStep.
        if (++num_consecutive_invalid_steps >= opt.max_num_consecutive_invalid_steps) {
          S.termination_type = kFailure;
          Format(S.message, sizeof S.message, "Number of consecutive invalid steps more than Solver::Options::max_num_consecutive_invalid_steps: %.0f", opt.max_num_consecutive_invalid_steps);
          break;
        }
        radius *= 0.5; reuse_diagonal = true;  // LevenbergMarquardtStrategy::StepIsInvalid
        it.cost = x_cost + S.fixed_cost; it.cost_change = 0.0; it.step_norm = 0.0; it.relative_decrease = 0.0;
        continue;
      }
      // ComputeCandidatePointAndEvaluateCost.
      Plus(x, delta, candidate_x);
      SetState(candidate_x);
      if (!Evaluate(&candidate_cost, false)) candidate_cost = std::numeric_limits<double>::max();  // "Step failed to evaluate."
      // ParameterToleranceReached.
      double sn = 0; for (int i = 0; i < n_amb; ++i) { const double d = x[i] - candidate_x[i]; sn += d * d; }
      it.step_norm = std::sqrt(sn);
      if (it.step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) {
        SetState(x);
        S.termination_type = kConvergence;
        Format(S.message, sizeof S.message, "Parameter tolerance reached. Relative step_norm: %e <= %e.", it.step_norm / (x_norm + opt.parameter_tolerance), opt.parameter_tolerance);
        break;
      }
      // FunctionToleranceReached.
      it.cost_change = x_cost - candidate_cost;
      if (std::fabs(it.cost_change) <= opt.function_tolerance * x_cost) {
        SetState(x);
        S.termination_type = kConvergence;
        Format(S.message, sizeof S.message, "Function tolerance reached. |cost_change|/cost: %e <= %e", std::fabs(it.cost_change) / x_cost, opt.function_tolerance);
        break;
      }
      // IsStepSuccessful: TrustRegionStepEvaluator::StepQuality.
      if (candidate_cost >= std::numeric_limits<double>::max()) it.relative_decrease = std::numeric_limits<double>::lowest();
      else it.relative_decrease = std::max((x_cost - candidate_cost) / model_cost_change, (reference_cost - candidate_cost) / model_cost_change);
      if (it.relative_decrease > opt.min_relative_decrease) {
        // HandleSuccessfulStep.
        x = candidate_x;
        x_norm = 0; for (double v : x) x_norm += v * v; x_norm = std::sqrt(x_norm);
        if (!evaluate_gradient_and_jacobian(false)) {
          S.termination_type = kFailure; Format(S.message, sizeof S.message, "Residual and Jacobian evaluation failed.");
          break;
        }
        it.step_is_successful = 1;
        // LevenbergMarquardtStrategy::StepAccepted.
        radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * it.relative_decrease - 1.0, 3));
        radius = std::min(opt.max_trust_region_radius, radius);
        decrease_factor = 2.0; reuse_diagonal = false;
        reference_cost = x_cost;
        ++S.num_successful_steps;
      } else {
        SetState(x);
        it.step_is_successful = 0;
        it.cost = candidate_cost + S.fixed_cost;
        // LevenbergMarquardtStrategy::StepRejected.
        radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
        ++S.num_unsuccessful_steps;
      }
    }
    // solver.cc SetSummaryFinalCost: min over the recorded iterations.
    S.final_cost = S.initial_cost;
    for (const auto& e : L) S.final_cost = std::min(S.final_cost, e.cost);
    S.num_iterations = int(L.size());
    S.total_time = NowSeconds() - t_start;
  }
};

}  // namespace orc
