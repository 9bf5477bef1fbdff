// TEST INFRASTRUCTURE ONLY (see ho_math.h header).  PARITY UNPINNED.
//
// Flat sliding-window restatement of one Gauss-Newton / LM iteration as ceres::Solve would run it
// on the reference's problem (reference internal/hyper/optimizers/ceres/optimizer.cpp:38-54,
// 189-274): per-factor ExteroceptiveCost::Evaluate -> loss correction (Huber 0.5 pixel :226,
// ScaledLoss 1.6e-5 inertial :267-268) -> manifold projection (Manifold::PlusJacobian,
// wrapper.hpp:36-42) -> normal equations -> LM step -> Manifold::Plus.
// The linear solve eliminates landmarks first (Schur complement) and factors the dense reduced
// system; the reference uses SPARSE_NORMAL_CHOLESKY on the full system (optimizer.cpp:46) -- same
// step up to rounding.
//
// Compact per-factor output layout (shared with the CUDA path, DESIGN.md "HBM layout"):
//   pixel   : r[2], Jp[2][6k] (pose control-point tangents [theta|rho] x k), Jl[2][3]
//   bearing : same slot as a pixel factor, r = [angle, 0], second Jacobian row zero (Huber 1.6e-3, :204)
//   manifold: r[6], Jp[6][6k] (no loss, :250)
//   inertial: r[6], Jp[6][6k], wg[k_b], wa[k_b] (bias basis weights), Jg[6][2] (gravity tangent)
// Tangents: theta GLOBAL rotation vector (R <- Exp(theta) R), rho additive; gravity tangent is the
// Ceres SphereManifold<3> tangent.
#pragma once
#include <cstdint>
#include <vector>

#include "ho_evaluators.h"
#include "ho_manifolds.h"

namespace ho {

struct Window {
  int k = 4, K = 0;
  std::vector<double> knots;  // K x 8
  int k_b = 4, Kbg = 0, Kba = 0;
  std::vector<double> bg, ba;  // x 4
  double gravity[3] = {0, 0, -9.81};
  int C = 0;
  std::vector<double> cams;  // C x 15 (T_bs 7, intrinsics 4, distortion 4)
  double imu[37] = {0};      // T_bs 7, i_g 6, i_a 6, S_g 9, X_a 9
  int L = 0;
  std::vector<double> landmarks;  // L x 3
  // visual list: pixel factors [0, Np) followed by bearing factors [Np, Nv)
  int Nv = 0, Np = 0;
  std::vector<double> v_stamp, v_pixel, v_z;  // v_pixel: pixel, or bearing xy; v_z: bearing z
  std::vector<int> v_cam, v_lm;
  // manifold (pose) factors
  int Nm = 0, P = 0;
  std::vector<double> m_stamp, m_meas, pose_sensors;  // m_meas Nm x 7, pose_sensors P x 7
  std::vector<int> m_sensor, m_base;
  int Ni = 0;
  std::vector<double> i_stamp, i_meas;
  // index maps (bind)
  std::vector<int> v_base, i_base, i_bg_base, i_ba_base;
  std::vector<uint8_t> knot_const;
  int gravity_const = 0, bias_const = 0;
  double huber_pixel = 0.5, huber_bearing = 1.6e-3, imu_loss_scale = 1.6e-5;
  double huber_of(int f) const { return f < Np ? huber_pixel : huber_bearing; }
  int quirks = 0;
  Basis basis, bias_basis;
  // solver state
  double radius = 1e4, decrease_factor = 2.0;
};

// Knot base index of a stamp (a2/a4): j = largest index with stamps[j] <= t (what an ordered
// upper_bound on the elements' stamps yields, reference optimizer.cpp:289), base = j - (k-1)/2.
// Returns -1 when the k control points are not all inside [0, K).  stride = doubles per knot,
// stamp_off = position of the stamp inside a knot.
inline int segment_base(const double* knots, int stride, int stamp_off, int K, int k, double t) {
  if (K < k) return -1;
  const double t0 = knots[stamp_off], t1 = knots[stride + stamp_off];
  long j = (long)std::floor((t - t0) / (t1 - t0));
  if (j < 0) j = 0;
  if (j > K - 2) j = K - 2;
  while (j > 0 && knots[(size_t)j * stride + stamp_off] > t) --j;
  while (j < K - 2 && knots[(size_t)(j + 1) * stride + stamp_off] <= t) ++j;
  if (!(knots[(size_t)j * stride + stamp_off] <= t && t < knots[(size_t)(j + 1) * stride + stamp_off])) return -1;
  const int base = (int)j - (k - 1) / 2;
  if (base < 0 || base + k - 1 > K - 1) return -1;
  return base;
}

inline int window_bind(Window* w) {
  basis_init(&w->basis, w->k);
  basis_init(&w->bias_basis, w->k_b);
  w->v_base.resize(w->Nv); w->i_base.resize(w->Ni); w->i_bg_base.resize(w->Ni); w->i_ba_base.resize(w->Ni);
  int bad = 0;
  for (int f = 0; f < w->Nv; ++f) {
    w->v_base[f] = segment_base(w->knots.data(), 8, 7, w->K, w->k, w->v_stamp[f]);
    if (w->v_base[f] < 0 || w->v_cam[f] < 0 || w->v_cam[f] >= w->C || w->v_lm[f] < 0 || w->v_lm[f] >= w->L) ++bad;
  }
  for (int f = 0; f < w->Ni; ++f) {
    w->i_base[f] = segment_base(w->knots.data(), 8, 7, w->K, w->k, w->i_stamp[f]);
    w->i_bg_base[f] = segment_base(w->bg.data(), 4, 3, w->Kbg, w->k_b, w->i_stamp[f]);
    w->i_ba_base[f] = segment_base(w->ba.data(), 4, 3, w->Kba, w->k_b, w->i_stamp[f]);
    if (w->i_base[f] < 0 || w->i_bg_base[f] < 0 || w->i_ba_base[f] < 0) ++bad;
  }
  w->m_base.resize(w->Nm);
  for (int f = 0; f < w->Nm; ++f) {
    w->m_base[f] = segment_base(w->knots.data(), 8, 7, w->K, w->k, w->m_stamp[f]);
    if (w->m_base[f] < 0 || w->m_sensor[f] < 0 || w->m_sensor[f] >= w->P) ++bad;
  }
  if ((int)w->knot_const.size() != w->K) w->knot_const.assign(w->K, 0);
  return bad;
}

// d ambient / d theta for a unit quaternion under R <- Exp(theta) R : 1/2 * EigenQuaternion PlusJacobian.
inline void quat_dtheta(const double* q, double* P /*4x3*/) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double M[12] = {w, z, -y, -z, w, x, y, -x, w, -x, -y, -z};
  for (int i = 0; i < 12; ++i) P[i] = 0.5 * M[i];
}

// Reference-shaped Evaluate of pixel factor f + projection to the compact tangent layout.
inline void window_eval_pixel(const Window& w, int f, bool want_J, double* r, double* Jp, double* Jl) {
  const int k = w.k;
  const bool bearing = f >= w.Np;
  Factor fac; fac.kind = bearing ? kBearing : kPixel; fac.stamp = w.v_stamp[f]; fac.k = k;
  fac.measurement[0] = w.v_pixel[2 * f]; fac.measurement[1] = w.v_pixel[2 * f + 1];
  if (bearing) fac.measurement[2] = w.v_z[f];
  const int nr = bearing ? 1 : 2;
  Layout L; layout_update(fac, &L);
  const double* ptrs[kMaxBlocks];
  const int base = w.v_base[f];
  for (int m = 0; m < k; ++m) ptrs[m] = &w.knots[8 * (base + m)];
  const double* cam = &w.cams[15 * w.v_cam[f]];
  ptrs[k] = cam; ptrs[k + 1] = cam + 7; ptrs[k + 2] = cam + 11;
  ptrs[k + 3] = &w.landmarks[3 * w.v_lm[f]];
  r[1] = 0.0;
  if (!want_J) { cost_evaluate(fac, L, w.basis, w.bias_basis, ptrs, r, nullptr, w.quirks); return; }
  double jbuf[2 * 8 * kMaxOrder + 6];
  double* jptrs[kMaxBlocks];
  for (int b = 0; b < L.num_blocks; ++b) jptrs[b] = nullptr;
  for (int m = 0; m < k; ++m) jptrs[m] = jbuf + 16 * m;
  jptrs[k + 3] = jbuf + 16 * k;
  cost_evaluate(fac, L, w.basis, w.bias_basis, ptrs, r, jptrs, w.quirks);
  if (bearing) { for (int i = 0; i < 6 * k; ++i) Jp[6 * k + i] = 0.0; for (int i = 3; i < 6; ++i) Jl[i] = 0.0; }
  for (int m = 0; m < k; ++m) {
    double P[12];
    quat_dtheta(ptrs[m], P);
    const double* Ja = jptrs[m];  // nr x 8
    for (int row = 0; row < nr; ++row) {
      for (int c = 0; c < 3; ++c)
        Jp[row * 6 * k + 6 * m + c] = Ja[8 * row] * P[c] + Ja[8 * row + 1] * P[3 + c] + Ja[8 * row + 2] * P[6 + c] + Ja[8 * row + 3] * P[9 + c];
      for (int c = 0; c < 3; ++c) Jp[row * 6 * k + 6 * m + 3 + c] = Ja[8 * row + 4 + c];
    }
  }
  for (int i = 0; i < 3 * nr; ++i) Jl[i] = jptrs[k + 3][i];
}

// Manifold (pose) factor f: r[6], Jp[6][6k].
inline void window_eval_manifold(const Window& w, int f, bool want_J, double* r, double* Jp) {
  const int k = w.k;
  Factor fac; fac.kind = kManifold; fac.stamp = w.m_stamp[f]; fac.k = k;
  for (int i = 0; i < 7; ++i) fac.measurement[i] = w.m_meas[7 * f + i];
  Layout L; layout_update(fac, &L);
  const double* ptrs[kMaxBlocks];
  const int base = w.m_base[f];
  for (int m = 0; m < k; ++m) ptrs[m] = &w.knots[8 * (base + m)];
  ptrs[k] = &w.pose_sensors[7 * w.m_sensor[f]];
  if (!want_J) { cost_evaluate(fac, L, w.basis, w.bias_basis, ptrs, r, nullptr, w.quirks); return; }
  double jbuf[6 * 8 * kMaxOrder];
  double* jptrs[kMaxBlocks];
  for (int b = 0; b < L.num_blocks; ++b) jptrs[b] = nullptr;
  for (int m = 0; m < k; ++m) jptrs[m] = jbuf + 48 * m;
  cost_evaluate(fac, L, w.basis, w.bias_basis, ptrs, r, jptrs, w.quirks);
  for (int m = 0; m < k; ++m) {
    double P[12];
    quat_dtheta(ptrs[m], P);
    const double* Ja = jptrs[m];  // 6 x 8
    for (int row = 0; row < 6; ++row) {
      for (int c = 0; c < 3; ++c)
        Jp[row * 6 * k + 6 * m + c] = Ja[8 * row] * P[c] + Ja[8 * row + 1] * P[3 + c] + Ja[8 * row + 2] * P[6 + c] + Ja[8 * row + 3] * P[9 + c];
      for (int c = 0; c < 3; ++c) Jp[row * 6 * k + 6 * m + 3 + c] = Ja[8 * row + 4 + c];
    }
  }
}

// J^T J / J^T r / cost of the manifold factors into a dense n x n (row-major) H and g.
inline double window_accumulate_manifold(const Window& w, double* H, double* g, int n) {
  const int nc = 6 * w.k;
  double cost = 0;
  for (int f = 0; f < w.Nm; ++f) {
    double r[6], Jp[6 * 6 * kMaxOrder];
    window_eval_manifold(w, f, true, r, Jp);
    const int c0 = 6 * w.m_base[f];
    for (int i = 0; i < 6; ++i) cost += 0.5 * r[i] * r[i];
    for (int a = 0; a < nc; ++a) {
      double ga = 0;
      for (int q = 0; q < 6; ++q) ga += Jp[q * nc + a] * r[q];
      g[c0 + a] += ga;
      for (int b = 0; b < nc; ++b) {
        double h = 0;
        for (int q = 0; q < 6; ++q) h += Jp[q * nc + a] * Jp[q * nc + b];
        H[(size_t)(c0 + a) * n + c0 + b] += h;
      }
    }
  }
  return cost;
}

inline void window_eval_inertial(const Window& w, int f, bool want_J, double* r, double* Jp, double* wg, double* wa, double* Jg) {
  const int k = w.k, kb = w.k_b;
  Factor fac; fac.kind = kInertial; fac.stamp = w.i_stamp[f]; fac.k = k; fac.k_bg = kb; fac.k_ba = kb;
  for (int i = 0; i < 6; ++i) fac.measurement[i] = w.i_meas[6 * f + i];
  Layout L; layout_update(fac, &L);
  const double* ptrs[kMaxBlocks];
  const int base = w.i_base[f];
  int n = 0;
  for (int m = 0; m < k; ++m) ptrs[n++] = &w.knots[8 * (base + m)];
  ptrs[n++] = w.imu; ptrs[n++] = w.imu + 7; ptrs[n++] = w.imu + 13; ptrs[n++] = w.imu + 19; ptrs[n++] = w.imu + 28;
  const int o_bg = n;
  for (int m = 0; m < kb; ++m) ptrs[n++] = &w.bg[4 * (w.i_bg_base[f] + m)];
  const int o_ba = n;
  for (int m = 0; m < kb; ++m) ptrs[n++] = &w.ba[4 * (w.i_ba_base[f] + m)];
  const int o_g = n;
  ptrs[n++] = w.gravity;
  if (!want_J) { cost_evaluate(fac, L, w.basis, w.bias_basis, ptrs, r, nullptr, w.quirks); return; }
  std::vector<double> jbuf((size_t)6 * L.num_parameters);
  double* jptrs[kMaxBlocks];
  for (int b = 0; b < L.num_blocks; ++b) jptrs[b] = nullptr;
  for (int m = 0; m < k; ++m) jptrs[m] = jbuf.data() + 6 * L.offsets[m];
  for (int m = 0; m < kb; ++m) { jptrs[o_bg + m] = jbuf.data() + 6 * L.offsets[o_bg + m]; jptrs[o_ba + m] = jbuf.data() + 6 * L.offsets[o_ba + m]; }
  jptrs[o_g] = jbuf.data() + 6 * L.offsets[o_g];
  cost_evaluate(fac, L, w.basis, w.bias_basis, ptrs, r, jptrs, w.quirks);
  for (int m = 0; m < k; ++m) {
    double P[12];
    quat_dtheta(ptrs[m], P);
    const double* Ja = jptrs[m];  // 6 x 8
    for (int row = 0; row < 6; ++row) {
      for (int c = 0; c < 3; ++c)
        Jp[row * 6 * k + 6 * m + c] = Ja[8 * row] * P[c] + Ja[8 * row + 1] * P[3 + c] + Ja[8 * row + 2] * P[6 + c] + Ja[8 * row + 3] * P[9 + c];
      for (int c = 0; c < 3; ++c) Jp[row * 6 * k + 6 * m + 3 + c] = Ja[8 * row + 4 + c];
    }
  }
  for (int m = 0; m < kb; ++m) { wg[m] = jptrs[o_bg + m][0]; wa[m] = jptrs[o_ba + m][3 * 4 + 0]; }
  Manifold sph = make_sphere(3, false);
  double PJ[6];
  manifold_plus_jacobian(sph, w.gravity, PJ);  // 3 x 2
  mat_mul(jptrs[o_g], PJ, Jg, 6, 3, 2);
}

// Reduced-system dof layout: [6K pose | 3Kbg | 3Kba | 2 gravity].
inline int reduced_size(const Window& w) { return 6 * w.K + 3 * w.Kbg + 3 * w.Kba + 2; }

// ceres::TrustRegionMinimizer termination tests (Ceres 2.x trust_region_minimizer.cc: IterationZero,
// ParameterToleranceReached, FunctionToleranceReached, HandleSuccessfulStep, FinalizeIterationAndCheckIfMinimizerCanContinue)
// restated from the public sources -- PARITY UNPINNED like the rest of the LM loop.  The reference leaves the
// tolerances at Ceres' defaults (reference internal/hyper/optimizers/ceres/optimizer.cpp:38-54).
struct Termination {
  double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8, min_radius = 1e-32;
  int type = 0;            // 0 running, 1 function, 2 parameter, 3 gradient tolerance, 4 minimum radius, 5 invalid steps
  int invalid_steps = 0;
  double gradient_max_norm = 0, step_norm = 0, x_norm = 0;
};

struct IterationOutput {
  std::vector<double> S, b;        // reduced system (n x n row-major full symmetric, n)
  std::vector<double> delta_p;     // n
  std::vector<double> delta_l;     // 3L
  double cost = 0, cost_new = 0, model_change = 0, rho = 0, radius = 0;
  int accepted = 0, spd = 1;
};

inline double huber_weight(double s, double delta, double* rho_s) {
  if (s <= delta * delta) { *rho_s = s; return 1.0; }
  const double rt = std::sqrt(s);
  *rho_s = 2.0 * delta * rt - delta * delta;
  return delta / rt;
}

inline double window_cost(const Window& w) {
  double cost = 0;
#pragma omp parallel for reduction(+ : cost) schedule(static)
  for (int f = 0; f < w.Nv; ++f) {
    double r[2], rho;
    window_eval_pixel(w, f, false, r, nullptr, nullptr);
    huber_weight(r[0] * r[0] + r[1] * r[1], w.huber_of(f), &rho);
    cost += 0.5 * rho;
  }
  for (int f = 0; f < w.Nm; ++f) {
    double r[6];
    window_eval_manifold(w, f, false, r, nullptr);
    for (int i = 0; i < 6; ++i) cost += 0.5 * r[i] * r[i];
  }
#pragma omp parallel for reduction(+ : cost) schedule(static)
  for (int f = 0; f < w.Ni; ++f) {
    double r[6], s = 0;
    window_eval_inertial(w, f, false, r, nullptr, nullptr, nullptr, nullptr);
    for (int i = 0; i < 6; ++i) s += r[i] * r[i];
    cost += 0.5 * w.imu_loss_scale * s;
  }
  return cost;
}

// In-place dense Cholesky (lower), returns false if not SPD.
inline bool cholesky(std::vector<double>& A, int n) {
  for (int j = 0; j < n; ++j) {
    double d = A[(size_t)j * n + j];
    for (int p = 0; p < j; ++p) d -= A[(size_t)j * n + p] * A[(size_t)j * n + p];
    if (!(d > 0)) return false;
    d = std::sqrt(d);
    A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[(size_t)i * n + j];
      for (int p = 0; p < j; ++p) s -= A[(size_t)i * n + p] * A[(size_t)j * n + p];
      A[(size_t)i * n + j] = s / d;
    }
  }
  return true;
}
inline void cholesky_solve(const std::vector<double>& Lm, int n, std::vector<double>& x) {
  for (int i = 0; i < n; ++i) {
    double s = x[i];
    for (int p = 0; p < i; ++p) s -= Lm[(size_t)i * n + p] * x[p];
    x[i] = s / Lm[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = x[i];
    for (int p = i + 1; p < n; ++p) s -= Lm[(size_t)p * n + i] * x[p];
    x[i] = s / Lm[(size_t)i * n + i];
  }
}
inline bool inv3_sym(const double* A, double* Ai) {
  const double a = A[0], b = A[1], c = A[2], d = A[4], e = A[5], f = A[8];
  const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
  const double det = a * c00 + b * c01 + c * c02;
  if (!(det > 0)) return false;
  const double id = 1.0 / det;
  Ai[0] = c00 * id; Ai[1] = c01 * id; Ai[2] = c02 * id;
  Ai[3] = Ai[1]; Ai[4] = (a * f - c * c) * id; Ai[5] = (b * c - a * e) * id;
  Ai[6] = Ai[2]; Ai[7] = Ai[5]; Ai[8] = (a * d - b * b) * id;
  return true;
}

inline void retract(Window* w, const std::vector<double>& dp, const std::vector<double>& dl) {
  for (int j = 0; j < w->K; ++j) {
    double* kn = &w->knots[8 * j];
    double qe[4], qn[4];
    quat_exp(&dp[6 * j], qe);
    quat_mul(qe, kn, qn);
    for (int c = 0; c < 4; ++c) kn[c] = qn[c];
    for (int c = 0; c < 3; ++c) kn[4 + c] += dp[6 * j + 3 + c];
  }
  int off = 6 * w->K;
  for (int j = 0; j < w->Kbg; ++j) for (int c = 0; c < 3; ++c) w->bg[4 * j + c] += dp[off + 3 * j + c];
  off += 3 * w->Kbg;
  for (int j = 0; j < w->Kba; ++j) for (int c = 0; c < 3; ++c) w->ba[4 * j + c] += dp[off + 3 * j + c];
  off += 3 * w->Kba;
  Manifold sph = make_sphere(3, false);
  double gn[3];
  manifold_plus(sph, w->gravity, &dp[off], gn);
  for (int c = 0; c < 3; ++c) w->gravity[c] = gn[c];
  for (int l = 0; l < w->L; ++l) for (int c = 0; c < 3; ++c) w->landmarks[3 * l + c] += dl[3 * l + c];
}

// One LM iteration.  If apply == 0 the state is left untouched (used for parity of S, b, delta).
inline void retract(Window* w, const std::vector<double>& dp, const std::vector<double>& dl);
inline void window_iterate(Window* w, IterationOutput* out, int apply, Termination* term = nullptr) {
  const int k = w->k, kb = w->k_b, K = w->K, n = reduced_size(*w), L = w->L;
  const int o_bg = 6 * K, o_ba = o_bg + 3 * w->Kbg, o_g = o_ba + 3 * w->Kba;
  std::vector<double> H((size_t)n * n, 0.0), g(n, 0.0);
  std::vector<double> V((size_t)9 * L, 0.0), gl((size_t)3 * L, 0.0);
  // W_l stored densely per landmark would be n x 3; keep per-observation contributions instead.
  struct Obs { int lm; int base; double Jp[2 * 6 * kMaxOrder]; double Jl[6]; double r[2]; };
  std::vector<Obs> obs(w->Nv);
  double cost = 0;
#pragma omp parallel for reduction(+ : cost) schedule(static)
  for (int f = 0; f < w->Nv; ++f) {
    Obs& o = obs[f];
    o.lm = w->v_lm[f]; o.base = w->v_base[f];
    window_eval_pixel(*w, f, true, o.r, o.Jp, o.Jl);
    double rho;
    const double wgt = huber_weight(o.r[0] * o.r[0] + o.r[1] * o.r[1], w->huber_of(f), &rho);
    cost += 0.5 * rho;
    const double sw = std::sqrt(wgt);
    for (int i = 0; i < 2 * 6 * k; ++i) o.Jp[i] *= sw;
    for (int i = 0; i < 6; ++i) o.Jl[i] *= sw;
    o.r[0] *= sw; o.r[1] *= sw;
  }
  for (int f = 0; f < w->Nv; ++f) {
    const Obs& o = obs[f];
    const int c0 = 6 * o.base, nc = 6 * k;
    for (int a = 0; a < nc; ++a) {
      g[c0 + a] += o.Jp[a] * o.r[0] + o.Jp[nc + a] * o.r[1];
      for (int bcol = 0; bcol < nc; ++bcol) H[(size_t)(c0 + a) * n + c0 + bcol] += o.Jp[a] * o.Jp[bcol] + o.Jp[nc + a] * o.Jp[nc + bcol];
    }
    for (int a = 0; a < 3; ++a) {
      gl[3 * o.lm + a] += o.Jl[a] * o.r[0] + o.Jl[3 + a] * o.r[1];
      for (int bcol = 0; bcol < 3; ++bcol) V[9 * o.lm + 3 * a + bcol] += o.Jl[a] * o.Jl[bcol] + o.Jl[3 + a] * o.Jl[3 + bcol];
    }
  }
  {
    struct IObs { double r[6]; double Jp[6 * 6 * kMaxOrder]; double wg[kMaxOrder], wa[kMaxOrder], Jg[12]; };
    std::vector<IObs> iobs(w->Ni);
#pragma omp parallel for reduction(+ : cost) schedule(static)
    for (int f = 0; f < w->Ni; ++f) {
      IObs& o = iobs[f];
      window_eval_inertial(*w, f, true, o.r, o.Jp, o.wg, o.wa, o.Jg);
      double s = 0;
      for (int i = 0; i < 6; ++i) s += o.r[i] * o.r[i];
      cost += 0.5 * w->imu_loss_scale * s;
    }
    const double sw = std::sqrt(w->imu_loss_scale);
    const int ncol = 6 * k + 6 * kb + 2;
    std::vector<double> J((size_t)6 * ncol);
    std::vector<int> idx(ncol);
    for (int f = 0; f < w->Ni; ++f) {
      const IObs& o = iobs[f];
      std::fill(J.begin(), J.end(), 0.0);
      int c = 0;
      for (int a = 0; a < 6 * k; ++a, ++c) { idx[c] = 6 * w->i_base[f] + a; for (int r = 0; r < 6; ++r) J[(size_t)r * ncol + c] = sw * o.Jp[r * 6 * k + a]; }
      for (int m = 0; m < kb; ++m) for (int a = 0; a < 3; ++a, ++c) { idx[c] = o_bg + 3 * (w->i_bg_base[f] + m) + a; J[(size_t)a * ncol + c] = sw * o.wg[m]; }
      for (int m = 0; m < kb; ++m) for (int a = 0; a < 3; ++a, ++c) { idx[c] = o_ba + 3 * (w->i_ba_base[f] + m) + a; J[(size_t)(3 + a) * ncol + c] = sw * o.wa[m]; }
      for (int a = 0; a < 2; ++a, ++c) { idx[c] = o_g + a; for (int r = 0; r < 6; ++r) J[(size_t)r * ncol + c] = sw * o.Jg[2 * r + a]; }
      for (int a = 0; a < ncol; ++a) {
        double ga = 0;
        for (int r = 0; r < 6; ++r) ga += J[(size_t)r * ncol + a] * sw * o.r[r];
        g[idx[a]] += ga;
        for (int bcol = 0; bcol < ncol; ++bcol) {
          double h = 0;
          for (int r = 0; r < 6; ++r) h += J[(size_t)r * ncol + a] * J[(size_t)r * ncol + bcol];
          H[(size_t)idx[a] * n + idx[bcol]] += h;
        }
      }
    }
  }
  cost += window_accumulate_manifold(*w, H.data(), g.data(), n);
  out->cost = cost;
  // Constant dofs (reference optimizer.cpp:322-328 SetParameterBlockConstant; setGravityConstant).
  std::vector<uint8_t> fixed(n, 0);
  for (int j = 0; j < K; ++j) if (w->knot_const[j]) for (int a = 0; a < 6; ++a) fixed[6 * j + a] = 1;
  if (w->bias_const) for (int a = o_bg; a < o_g; ++a) fixed[a] = 1;
  if (w->gravity_const) { fixed[o_g] = fixed[o_g + 1] = 1; }
  // LM damping: mu * clamp(diag(H)) (Ceres LevenbergMarquardtStrategy, min/max_lm_diagonal 1e-6/1e32).
  const double mu = 1.0 / w->radius;
  std::vector<double> Dp(n), Dl((size_t)3 * L);
  for (int a = 0; a < n; ++a) Dp[a] = std::min(std::max(H[(size_t)a * n + a], 1e-6), 1e32);
  for (int l = 0; l < L; ++l) for (int a = 0; a < 3; ++a) Dl[3 * l + a] = std::min(std::max(V[9 * l + 4 * a], 1e-6), 1e32);
  std::vector<double> S = H, b(n);
  for (int a = 0; a < n; ++a) { S[(size_t)a * n + a] += mu * Dp[a]; b[a] = -g[a]; }
  // Schur complement over landmarks.
  std::vector<double> Vi((size_t)9 * L, 0.0);
  std::vector<std::vector<int>> lm_obs(L);
  for (int f = 0; f < w->Nv; ++f) lm_obs[obs[f].lm].push_back(f);
  for (int l = 0; l < L; ++l) {
    if (lm_obs[l].empty()) continue;
    double Vd[9];
    for (int i = 0; i < 9; ++i) Vd[i] = V[9 * l + i];
    for (int a = 0; a < 3; ++a) Vd[4 * a] += mu * Dl[3 * l + a];
    if (!inv3_sym(Vd, &Vi[9 * l])) { out->spd = 0; continue; }
    // W_l rows: for each obs, rows 6*base..6*base+6k: W = Jp^T Jl (6k x 3)
    const int nc = 6 * k;
    std::vector<double> Wl((size_t)n * 3, 0.0);
    std::vector<int> rows;
    for (int f : lm_obs[l]) {
      const Obs& o = obs[f];
      for (int a = 0; a < nc; ++a)
        for (int c = 0; c < 3; ++c) Wl[(size_t)(6 * o.base + a) * 3 + c] += o.Jp[a] * o.Jl[c] + o.Jp[nc + a] * o.Jl[3 + c];
    }
    for (int a = 0; a < n; ++a) if (Wl[3 * a] != 0 || Wl[3 * a + 1] != 0 || Wl[3 * a + 2] != 0) rows.push_back(a);
    const double* Vil = &Vi[9 * l];
    for (int a : rows) {
      double WV[3];
      for (int c = 0; c < 3; ++c) WV[c] = Wl[3 * a] * Vil[c] + Wl[3 * a + 1] * Vil[3 + c] + Wl[3 * a + 2] * Vil[6 + c];
      b[a] += WV[0] * gl[3 * l] + WV[1] * gl[3 * l + 1] + WV[2] * gl[3 * l + 2];
      for (int bb : rows) S[(size_t)a * n + bb] -= WV[0] * Wl[3 * bb] + WV[1] * Wl[3 * bb + 1] + WV[2] * Wl[3 * bb + 2];
    }
  }
  for (int a = 0; a < n; ++a)
    if (fixed[a]) {
      for (int c = 0; c < n; ++c) { S[(size_t)a * n + c] = 0; S[(size_t)c * n + a] = 0; }
      S[(size_t)a * n + a] = 1.0; b[a] = 0.0;
    }
  out->S = S; out->b = b;
  if (term) {
    // |x - Plus(x, -g)|_inf at the linearisation point, in Ceres' local coordinates (its rotation coordinate is half of
    // this restatement's theta: Plus(x, -g_ceres) rotates by -4 g_theta); fixed dofs do not move.
    std::vector<double> dpg(n), dlg((size_t)3 * L);
    for (int a = 0; a < n; ++a) dpg[a] = fixed[a] ? 0.0 : -g[a];
    for (int j = 0; j < K; ++j) for (int c = 0; c < 3; ++c) dpg[6 * j + c] *= 4.0;
    for (int i = 0; i < 3 * L; ++i) dlg[i] = -gl[i];
    Window xp = *w;
    retract(&xp, dpg, dlg);
    double gmax = 0;
    for (size_t i = 0; i < w->knots.size(); ++i) gmax = std::max(gmax, std::fabs(xp.knots[i] - w->knots[i]));
    for (size_t i = 0; i < w->bg.size(); ++i) gmax = std::max(gmax, std::fabs(xp.bg[i] - w->bg[i]));
    for (size_t i = 0; i < w->ba.size(); ++i) gmax = std::max(gmax, std::fabs(xp.ba[i] - w->ba[i]));
    for (int c = 0; c < 3; ++c) gmax = std::max(gmax, std::fabs(xp.gravity[c] - w->gravity[c]));
    for (int l = 0; l < L; ++l) if (!lm_obs[l].empty()) for (int c = 0; c < 3; ++c) gmax = std::max(gmax, std::fabs(gl[3 * l + c]));
    term->gradient_max_norm = gmax;
    if (term->gradient_tolerance > 0 && gmax <= term->gradient_tolerance) { term->type = 3; return; }
  }
  std::vector<double> Lc = S, dp = b;
  if (!cholesky(Lc, n)) {
    out->spd = 0; out->delta_p.assign(n, 0.0); out->delta_l.assign((size_t)3 * L, 0.0);
    if (apply) {   // HandleInvalidStep: the strategy shrinks the radius as for a rejected step
      w->radius /= w->decrease_factor; w->decrease_factor *= 2.0; out->radius = w->radius;
      if (term && ++term->invalid_steps >= 5) term->type = 5;
    }
    return;
  }
  if (term) term->invalid_steps = 0;
  cholesky_solve(Lc, n, dp);
  // Back-substitution: dl = V^{-1} (-g_l - W^T dp)
  std::vector<double> dl((size_t)3 * L, 0.0), rhs((size_t)3 * L);
  for (int i = 0; i < 3 * L; ++i) rhs[i] = -gl[i];
  const int nc = 6 * k;
  for (int f = 0; f < w->Nv; ++f) {
    const Obs& o = obs[f];
    double t0 = 0, t1 = 0;
    for (int a = 0; a < nc; ++a) { t0 += o.Jp[a] * dp[6 * o.base + a]; t1 += o.Jp[nc + a] * dp[6 * o.base + a]; }
    for (int c = 0; c < 3; ++c) rhs[3 * o.lm + c] -= o.Jl[c] * t0 + o.Jl[3 + c] * t1;
  }
  for (int l = 0; l < L; ++l) {
    if (lm_obs[l].empty()) continue;
    m3_vec(&Vi[9 * l], &rhs[3 * l], &dl[3 * l]);
  }
  out->delta_p = dp; out->delta_l = dl;
  // model_cost_change = -delta^T g - 1/2 delta^T H delta = 1/2 (-delta^T g + mu delta^T D delta)  (free dofs)
  double dg = 0, dDd = 0;
  for (int a = 0; a < n; ++a) if (!fixed[a]) { dg += dp[a] * g[a]; dDd += dp[a] * dp[a] * Dp[a]; }
  for (int i = 0; i < 3 * L; ++i) { dg += dl[i] * gl[i]; dDd += dl[i] * dl[i] * Dl[i]; }
  out->model_change = 0.5 * (-dg + mu * dDd);
  if (!apply) return;
  Window trial = *w;
  retract(&trial, dp, dl);
  out->cost_new = window_cost(trial);
  out->rho = (out->cost - out->cost_new) / out->model_change;
  if (term) {
    // ambient norms over the non-constant parameter blocks (Stamped blocks count their stamp coordinate)
    double s2 = 0, x2 = 0;
    auto acc = [&](const double* t, const double* x, int cnt) { for (int c = 0; c < cnt; ++c) { s2 += (t[c] - x[c]) * (t[c] - x[c]); x2 += x[c] * x[c]; } };
    for (int j = 0; j < K; ++j) if (!w->knot_const[j]) acc(&trial.knots[8 * j], &w->knots[8 * j], 8);
    if (!w->bias_const) {
      for (int j = 0; j < w->Kbg; ++j) acc(&trial.bg[4 * j], &w->bg[4 * j], 4);
      for (int j = 0; j < w->Kba; ++j) acc(&trial.ba[4 * j], &w->ba[4 * j], 4);
    }
    if (!w->gravity_const) acc(trial.gravity, w->gravity, 3);
    for (int l = 0; l < L; ++l) if (!lm_obs[l].empty()) acc(&trial.landmarks[3 * l], &w->landmarks[3 * l], 3);
    term->step_norm = std::sqrt(s2); term->x_norm = std::sqrt(x2);
    if (term->parameter_tolerance > 0 && term->step_norm <= term->parameter_tolerance * (term->x_norm + term->parameter_tolerance)) { term->type = 2; return; }
    if (term->function_tolerance > 0 && std::fabs(out->cost - out->cost_new) <= term->function_tolerance * out->cost) { term->type = 1; return; }
  }
  // Ceres TrustRegionMinimizer / LevenbergMarquardtStrategy step acceptance (min_relative_decrease 1e-3).
  if (out->model_change > 0 && out->rho > 1e-3) {
    out->accepted = 1;
    const double radius = w->radius;
    w->knots = trial.knots; w->bg = trial.bg; w->ba = trial.ba; w->landmarks = trial.landmarks;
    for (int c = 0; c < 3; ++c) w->gravity[c] = trial.gravity[c];
    const double t = 2.0 * out->rho - 1.0;
    w->radius = std::min(1e16, radius / std::max(1.0 / 3.0, 1.0 - t * t * t));
    w->decrease_factor = 2.0;
  } else {
    out->accepted = 0;
    w->radius /= w->decrease_factor;
    w->decrease_factor *= 2.0;
  }
  out->radius = w->radius;
  if (term && term->min_radius > 0 && w->radius <= term->min_radius) term->type = 4;
}

// ceres::Solve with max_num_iterations = max_iter: iterations until a termination test fires.  Returns the number performed.
inline int window_optimize(Window* w, int max_iter, Termination* term, std::vector<IterationOutput>* outs) {
  int performed = 0;
  for (int it = 0; it < max_iter && term->type == 0; ++it) {
    IterationOutput out;
    window_iterate(w, &out, 1, term);
    if (term->type == 1 || term->type == 2 || term->type == 3) break;   // ended before this iteration's step was applied / counted
    outs->push_back(out);
    performed += 1;
  }
  return performed;
}


// ---------------------------------------------------------------------------------------------
// Sharded form of the normal equations (multi-GPU layout, DESIGN.md "Multi-GPU"): each rank builds
//   packed = [ S (n*n) | b (n) | diagH (n) | g (n) | cost | pad ]
// from ITS factors only -- S = H_pp - sum_l W_l (V_l + mu D_l)^-1 W_l^T without pose damping,
// b = -g + sum_l W_l (V_l + mu D_l)^-1 g_l -- the packed buffers are summed across ranks, and
// packed_finalize() then applies the pose damping mu * clamp(diagH) and the constant-dof mask.
// Exact (not approximate) as long as all observations of a landmark live on one rank.
// ---------------------------------------------------------------------------------------------
inline void window_build_packed(const Window& w, std::vector<double>* packed) {
  const int k = w.k, kb = w.k_b, K = w.K, n = reduced_size(w), L = w.L;
  const int o_bg = 6 * K, o_ba = o_bg + 3 * w.Kbg, o_g = o_ba + 3 * w.Kba;
  packed->assign((size_t)n * n + 3 * (size_t)n + 2, 0.0);
  double* S = packed->data();
  double* b = S + (size_t)n * n;
  double* diagH = b + n;
  double* g = diagH + n;
  double* cost = g + n;
  std::vector<double> V((size_t)9 * L, 0.0), gl((size_t)3 * L, 0.0);
  struct Obs { int lm; int base; double Jp[2 * 6 * kMaxOrder]; double Jl[6]; double r[2]; };
  std::vector<Obs> obs(w.Nv);
  const int nc = 6 * k;
  for (int f = 0; f < w.Nv; ++f) {
    Obs& o = obs[f];
    o.lm = w.v_lm[f]; o.base = w.v_base[f];
    window_eval_pixel(w, f, true, o.r, o.Jp, o.Jl);
    double rho;
    const double wgt = huber_weight(o.r[0] * o.r[0] + o.r[1] * o.r[1], w.huber_of(f), &rho);
    *cost += 0.5 * rho;
    const double sw = std::sqrt(wgt);
    for (int i = 0; i < 2 * nc; ++i) o.Jp[i] *= sw;
    for (int i = 0; i < 6; ++i) o.Jl[i] *= sw;
    o.r[0] *= sw; o.r[1] *= sw;
    const int c0 = 6 * o.base;
    for (int a = 0; a < nc; ++a) {
      g[c0 + a] += o.Jp[a] * o.r[0] + o.Jp[nc + a] * o.r[1];
      for (int bc = 0; bc < nc; ++bc) S[(size_t)(c0 + a) * n + c0 + bc] += o.Jp[a] * o.Jp[bc] + o.Jp[nc + a] * o.Jp[nc + bc];
    }
    for (int a = 0; a < 3; ++a) {
      gl[3 * o.lm + a] += o.Jl[a] * o.r[0] + o.Jl[3 + a] * o.r[1];
      for (int bc = 0; bc < 3; ++bc) V[9 * o.lm + 3 * a + bc] += o.Jl[a] * o.Jl[bc] + o.Jl[3 + a] * o.Jl[3 + bc];
    }
  }
  {
    const double sw = std::sqrt(w.imu_loss_scale);
    const int ncol = 6 * k + 6 * kb + 2;
    std::vector<double> J((size_t)6 * ncol);
    std::vector<int> idx(ncol);
    for (int f = 0; f < w.Ni; ++f) {
      double r[6], Jp[6 * 6 * kMaxOrder], wg[kMaxOrder], wa[kMaxOrder], Jg[12];
      window_eval_inertial(w, f, true, r, Jp, wg, wa, Jg);
      double s2 = 0;
      for (int i = 0; i < 6; ++i) s2 += r[i] * r[i];
      *cost += 0.5 * w.imu_loss_scale * s2;
      std::fill(J.begin(), J.end(), 0.0);
      int c = 0;
      for (int a = 0; a < 6 * k; ++a, ++c) { idx[c] = 6 * w.i_base[f] + a; for (int rr = 0; rr < 6; ++rr) J[(size_t)rr * ncol + c] = sw * Jp[rr * 6 * k + a]; }
      for (int m = 0; m < kb; ++m) for (int a = 0; a < 3; ++a, ++c) { idx[c] = o_bg + 3 * (w.i_bg_base[f] + m) + a; J[(size_t)a * ncol + c] = sw * wg[m]; }
      for (int m = 0; m < kb; ++m) for (int a = 0; a < 3; ++a, ++c) { idx[c] = o_ba + 3 * (w.i_ba_base[f] + m) + a; J[(size_t)(3 + a) * ncol + c] = sw * wa[m]; }
      for (int a = 0; a < 2; ++a, ++c) { idx[c] = o_g + a; for (int rr = 0; rr < 6; ++rr) J[(size_t)rr * ncol + c] = sw * Jg[2 * rr + a]; }
      for (int a = 0; a < ncol; ++a) {
        double ga = 0;
        for (int rr = 0; rr < 6; ++rr) ga += J[(size_t)rr * ncol + a] * sw * r[rr];
        g[idx[a]] += ga;
        for (int bc = 0; bc < ncol; ++bc) {
          double hh = 0;
          for (int rr = 0; rr < 6; ++rr) hh += J[(size_t)rr * ncol + a] * J[(size_t)rr * ncol + bc];
          S[(size_t)idx[a] * n + idx[bc]] += hh;
        }
      }
    }
  }
  *cost += window_accumulate_manifold(w, S, g, n);
  for (int a = 0; a < n; ++a) { diagH[a] = S[(size_t)a * n + a]; b[a] = -g[a]; }
  const double mu = 1.0 / w.radius;
  std::vector<std::vector<int>> lm_obs(L);
  for (int f = 0; f < w.Nv; ++f) lm_obs[obs[f].lm].push_back(f);
  for (int l = 0; l < L; ++l) {
    if (lm_obs[l].empty()) continue;
    double Vd[9], Vi[9];
    for (int i = 0; i < 9; ++i) Vd[i] = V[9 * l + i];
    for (int a = 0; a < 3; ++a) Vd[4 * a] += mu * std::min(std::max(V[9 * l + 4 * a], 1e-6), 1e32);
    if (!inv3_sym(Vd, Vi)) continue;
    std::vector<double> Wl((size_t)n * 3, 0.0);
    for (int f : lm_obs[l]) {
      const Obs& o = obs[f];
      for (int a = 0; a < nc; ++a)
        for (int c = 0; c < 3; ++c) Wl[(size_t)(6 * o.base + a) * 3 + c] += o.Jp[a] * o.Jl[c] + o.Jp[nc + a] * o.Jl[3 + c];
    }
    std::vector<int> rows;
    for (int a = 0; a < n; ++a) if (Wl[3 * a] != 0 || Wl[3 * a + 1] != 0 || Wl[3 * a + 2] != 0) rows.push_back(a);
    for (int a : rows) {
      double WV[3];
      for (int c = 0; c < 3; ++c) WV[c] = Wl[3 * a] * Vi[c] + Wl[3 * a + 1] * Vi[3 + c] + Wl[3 * a + 2] * Vi[6 + c];
      b[a] += WV[0] * gl[3 * l] + WV[1] * gl[3 * l + 1] + WV[2] * gl[3 * l + 2];
      for (int bb : rows) S[(size_t)a * n + bb] -= WV[0] * Wl[3 * bb] + WV[1] * Wl[3 * bb + 1] + WV[2] * Wl[3 * bb + 2];
    }
  }
}

inline void packed_finalize(const Window& w, std::vector<double>* packed) {
  const int n = reduced_size(w), K = w.K;
  const int o_bg = 6 * K, o_g = o_bg + 3 * w.Kbg + 3 * w.Kba;
  double* S = packed->data();
  double* b = S + (size_t)n * n;
  const double* diagH = b + n;
  const double mu = 1.0 / w.radius;
  std::vector<uint8_t> fixed(n, 0);
  for (int j = 0; j < K; ++j) if (w.knot_const[j]) for (int a = 0; a < 6; ++a) fixed[6 * j + a] = 1;
  if (w.bias_const) for (int a = o_bg; a < o_g; ++a) fixed[a] = 1;
  if (w.gravity_const) { fixed[o_g] = fixed[o_g + 1] = 1; }
  for (int a = 0; a < n; ++a) S[(size_t)a * n + a] += mu * std::min(std::max(diagH[a], 1e-6), 1e32);
  for (int a = 0; a < n; ++a)
    if (fixed[a]) {
      for (int c = 0; c < n; ++c) { S[(size_t)a * n + c] = 0; S[(size_t)c * n + a] = 0; }
      S[(size_t)a * n + a] = 1.0; b[a] = 0.0;
    }
}

}  // namespace ho
