// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into, imported by or called from the
// product path (calico_b200/); only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use it.
//
// Pin status (tests/test_oracle_pins.py, tests/test_spline_fit.py; details in DESIGN.md §2): projections of OpenCv5 / OpenCv8 /
// KannalaBrandt against OpenCV golden vectors, spline basis / fit / derivatives, SO(3) and IMU kinematics against the reference tests'
// constants and tolerances, the trust-region radius schedule against the Ceres log stored in the reference's notebook, and the
// reference's integration test acceptance. PARITY UNPINNED: DoubleSphere / FieldOfView / UnifiedCamera / ExtendedUnifiedCamera
// projections (restated line by line, no independent golden source) and the per-iteration LM cost sequence (pinned to Ceres's
// published semantics, not to a Ceres binary: Ceres / Eigen / Abseil are absent, the reference cannot be compiled here).
//
// Forward-mode dual numbers: the same mathematical object as ceres::Jet<double, N>
// (Ceres is an un-vendored dependency of the reference: CMakeLists.txt:15, absent here).
// The reference's cost functors are templated on T and are differentiated by
// ceres::DynamicAutoDiffCostFunction in strides of 4 partials
// (camera_cost_functor.cpp:25, gyroscope_cost_functor.cpp:24, accelerometer_cost_functor.cpp:25).
#pragma once
#include <cmath>

namespace orc {

template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0.0) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; }  // NOLINT implicit
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; v[k] = 1.0; }
};

#define ORC_JET_LOOP for (int i = 0; i < N; ++i)

template <int N> inline Jet<N> operator+(const Jet<N>& f) { return f; }
template <int N> inline Jet<N> operator-(const Jet<N>& f) { Jet<N> r; r.a = -f.a; ORC_JET_LOOP r.v[i] = -f.v[i]; return r; }
template <int N> inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) { Jet<N> r; r.a = f.a + g.a; ORC_JET_LOOP r.v[i] = f.v[i] + g.v[i]; return r; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) { Jet<N> r; r.a = f.a - g.a; ORC_JET_LOOP r.v[i] = f.v[i] - g.v[i]; return r; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) { Jet<N> r; r.a = f.a * g.a; ORC_JET_LOOP r.v[i] = f.a * g.v[i] + f.v[i] * g.a; return r; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> r; const double gi = 1.0 / g.a; const double q = f.a * gi; r.a = q;
  ORC_JET_LOOP r.v[i] = (f.v[i] - q * g.v[i]) * gi; return r;
}
template <int N> inline Jet<N> operator+(const Jet<N>& f, double s) { Jet<N> r = f; r.a += s; return r; }
template <int N> inline Jet<N> operator+(double s, const Jet<N>& f) { Jet<N> r = f; r.a += s; return r; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, double s) { Jet<N> r = f; r.a -= s; return r; }
template <int N> inline Jet<N> operator-(double s, const Jet<N>& f) { Jet<N> r; r.a = s - f.a; ORC_JET_LOOP r.v[i] = -f.v[i]; return r; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, double s) { Jet<N> r; r.a = f.a * s; ORC_JET_LOOP r.v[i] = f.v[i] * s; return r; }
template <int N> inline Jet<N> operator*(double s, const Jet<N>& f) { return f * s; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, double s) { const double si = 1.0 / s; return f * si; }
template <int N> inline Jet<N> operator/(double s, const Jet<N>& g) {
  Jet<N> r; const double gi = 1.0 / g.a; r.a = s * gi; const double m = -s * gi * gi;
  ORC_JET_LOOP r.v[i] = m * g.v[i]; return r;
}
template <int N> inline Jet<N>& operator+=(Jet<N>& f, const Jet<N>& g) { f = f + g; return f; }
template <int N> inline Jet<N>& operator-=(Jet<N>& f, const Jet<N>& g) { f = f - g; return f; }
template <int N> inline Jet<N>& operator*=(Jet<N>& f, const Jet<N>& g) { f = f * g; return f; }
template <int N> inline Jet<N>& operator/=(Jet<N>& f, const Jet<N>& g) { f = f / g; return f; }
template <int N> inline Jet<N>& operator*=(Jet<N>& f, double s) { f = f * s; return f; }
template <int N> inline Jet<N>& operator+=(Jet<N>& f, double s) { f.a += s; return f; }

// Comparisons look at the scalar part only (as ceres::Jet does) — SURVEY §8 parity trap 3.
#define ORC_JET_CMP(op) \
  template <int N> inline bool operator op(const Jet<N>& f, const Jet<N>& g) { return f.a op g.a; } \
  template <int N> inline bool operator op(const Jet<N>& f, double g) { return f.a op g; }          \
  template <int N> inline bool operator op(double f, const Jet<N>& g) { return f op g.a; }
ORC_JET_CMP(<) ORC_JET_CMP(<=) ORC_JET_CMP(>) ORC_JET_CMP(>=) ORC_JET_CMP(==) ORC_JET_CMP(!=)
#undef ORC_JET_CMP

template <int N> inline Jet<N> sqrt(const Jet<N>& f) { Jet<N> r; r.a = std::sqrt(f.a); const double m = 0.5 / r.a; ORC_JET_LOOP r.v[i] = m * f.v[i]; return r; }
template <int N> inline Jet<N> sin(const Jet<N>& f) { Jet<N> r; r.a = std::sin(f.a); const double m = std::cos(f.a); ORC_JET_LOOP r.v[i] = m * f.v[i]; return r; }
template <int N> inline Jet<N> cos(const Jet<N>& f) { Jet<N> r; r.a = std::cos(f.a); const double m = -std::sin(f.a); ORC_JET_LOOP r.v[i] = m * f.v[i]; return r; }
template <int N> inline Jet<N> tan(const Jet<N>& f) { Jet<N> r; r.a = std::tan(f.a); const double m = 1.0 + r.a * r.a; ORC_JET_LOOP r.v[i] = m * f.v[i]; return r; }
template <int N> inline Jet<N> atan(const Jet<N>& f) { Jet<N> r; r.a = std::atan(f.a); const double m = 1.0 / (1.0 + f.a * f.a); ORC_JET_LOOP r.v[i] = m * f.v[i]; return r; }
template <int N> inline Jet<N> asin(const Jet<N>& f) { Jet<N> r; r.a = std::asin(f.a); const double m = 1.0 / std::sqrt(1.0 - f.a * f.a); ORC_JET_LOOP r.v[i] = m * f.v[i]; return r; }
template <int N> inline Jet<N> acos(const Jet<N>& f) { Jet<N> r; r.a = std::acos(f.a); const double m = -1.0 / std::sqrt(1.0 - f.a * f.a); ORC_JET_LOOP r.v[i] = m * f.v[i]; return r; }
#undef ORC_JET_LOOP

inline double sqrt(double x) { return std::sqrt(x); }
inline double sin(double x) { return std::sin(x); }
inline double cos(double x) { return std::cos(x); }
inline double tan(double x) { return std::tan(x); }
inline double atan(double x) { return std::atan(x); }
inline double asin(double x) { return std::asin(x); }
inline double acos(double x) { return std::acos(x); }

inline double scalar_part(double x) { return x; }
template <int N> inline double scalar_part(const Jet<N>& x) { return x.a; }

}  // namespace orc
