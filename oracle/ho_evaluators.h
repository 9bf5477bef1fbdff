// TEST INFRASTRUCTURE ONLY (see ho_math.h header).  PARITY UNPINNED for the HyperVariables /
// HyperSensors arithmetic (SE3 group ops, camera model, IMU intrinsics); the evaluator bodies
// follow the reference's in-tree code line by line:
//   VisualPixelEvaluator<SE3>::evaluate   reference internal/hyper/optimizers/evaluators/pixel.cpp:16-146
//   InertialEvaluator<SE3>::evaluate      reference internal/hyper/optimizers/evaluators/inertial.cpp:13-205
//   ExteroceptiveCost::update / Evaluate  reference internal/hyper/optimizers/ceres/costs/exteroceptive.cpp:25-160
#pragma once
#include <vector>

#include "ho_spline.h"

namespace ho {

// ---------------------------------------------------------------------------------------------
// SE3 = [q(4) | p(3)], tangent [theta(3) | rho(3)], R <- Exp(theta) R, p <- p + rho.
// Roles fixed by reference pixel.cpp:58-60: T_ws = T_wb.groupPlus(T_bs), p_s = T_sw.vectorPlus(p_w).
// Jacobians are 6x6 / 3x6 row-major, tangent -> tangent.
// ---------------------------------------------------------------------------------------------
inline void se3_group_plus(const double* A, const double* B, double* C, double* J_lhs, double* J_rhs) {
  double RA[9], RApB[3];
  quat_to_rot(A, RA);
  m3_vec(RA, B + 4, RApB);
  quat_mul(A, B, C);
  for (int c = 0; c < 3; ++c) C[4 + c] = A[4 + c] + RApB[c];
  if (J_lhs) {
    std::memset(J_lhs, 0, 36 * sizeof(double));
    double H[9];
    hat(RApB, H);
    for (int i = 0; i < 3; ++i) {
      J_lhs[6 * i + i] = 1.0;
      J_lhs[6 * (3 + i) + 3 + i] = 1.0;
      for (int j = 0; j < 3; ++j) J_lhs[6 * (3 + i) + j] = -H[3 * i + j];
    }
  }
  if (J_rhs) {
    std::memset(J_rhs, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        J_rhs[6 * i + j] = RA[3 * i + j];
        J_rhs[6 * (3 + i) + 3 + j] = RA[3 * i + j];
      }
  }
}

inline void se3_group_inverse(const double* A, double* C, double* J) {
  double RA[9], RT[9], t[3];
  quat_to_rot(A, RA);
  m3_transpose(RA, RT);
  m3_vec(RT, A + 4, t);
  quat_conj(A, C);
  for (int c = 0; c < 3; ++c) C[4 + c] = -t[c];
  if (J) {
    std::memset(J, 0, 36 * sizeof(double));
    double H[9], RTH[9];
    hat(A + 4, H);
    m3_mul(RT, H, RTH);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        J[6 * i + j] = -RT[3 * i + j];
        J[6 * (3 + i) + j] = -RTH[3 * i + j];
        J[6 * (3 + i) + 3 + j] = -RT[3 * i + j];
      }
  }
}

inline void se3_vector_plus(const double* A, const double* v, double* r, double* J /*3x6*/) {
  double RA[9], Rv[3];
  quat_to_rot(A, RA);
  m3_vec(RA, v, Rv);
  for (int c = 0; c < 3; ++c) r[c] = Rv[c] + A[4 + c];
  if (J) {
    double H[9];
    hat(Rv, H);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        J[6 * i + j] = -H[3 * i + j];
        J[6 * i + 3 + j] = (i == j) ? 1.0 : 0.0;
      }
  }
}

// SE3JacobianAdapter (reference pixel.cpp:141): d tangent / d ambient, 6x7.
inline void se3_adapter(const double* T, double* A /*6x7*/) {
  std::memset(A, 0, 42 * sizeof(double));
  double Aq[12];
  su2_adapter(T, Aq);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 4; ++c) A[7 * r + c] = Aq[4 * r + c];
    A[7 * (3 + r) + 4 + r] = 1.0;
  }
}

// ---------------------------------------------------------------------------------------------
// Camera model (reference call sites pixel.cpp:61-63,87,95,99).
// ---------------------------------------------------------------------------------------------
inline void project_to_plane(const double* p, double* n, double* J /*2x3*/) {
  const double iz = 1.0 / p[2];
  n[0] = p[0] * iz; n[1] = p[1] * iz;
  if (J) {
    J[0] = iz; J[1] = 0; J[2] = -p[0] * iz * iz;
    J[3] = 0; J[4] = iz; J[5] = -p[1] * iz * iz;
  }
}
// Radial-tangential, parameters [k1 k2 p1 p2] (reference settings.yaml:42-45).
inline void radtan_distort(const double* d, const double* n, double* o, double* J_n /*2x2*/, double* J_d /*2x4*/) {
  const double k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3];
  const double x = n[0], y = n[1];
  const double r2 = x * x + y * y;
  const double rad = 1.0 + k1 * r2 + k2 * r2 * r2;
  o[0] = x * rad + 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x);
  o[1] = y * rad + p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y;
  if (J_n) {
    const double g = k1 + 2.0 * k2 * r2;
    const double drx = 2.0 * x * g, dry = 2.0 * y * g;
    J_n[0] = rad + x * drx + 2.0 * p1 * y + 6.0 * p2 * x;
    J_n[1] = x * dry + 2.0 * p1 * x + 2.0 * p2 * y;
    J_n[2] = y * drx + 2.0 * p1 * x + 2.0 * p2 * y;
    J_n[3] = rad + y * dry + 6.0 * p1 * y + 2.0 * p2 * x;
  }
  if (J_d) {
    J_d[0] = x * r2; J_d[1] = x * r2 * r2; J_d[2] = 2.0 * x * y; J_d[3] = r2 + 2.0 * x * x;
    J_d[4] = y * r2; J_d[5] = y * r2 * r2; J_d[6] = r2 + 2.0 * y * y; J_d[7] = 2.0 * x * y;
  }
}
// Intrinsics [cx cy fx fy] (reference settings.yaml:38-40).
inline void denormalize(const double* in, const double* d, double* px, double* J_d /*2x2*/, double* J_i /*2x4*/) {
  px[0] = in[2] * d[0] + in[0];
  px[1] = in[3] * d[1] + in[1];
  if (J_d) { J_d[0] = in[2]; J_d[1] = 0; J_d[2] = 0; J_d[3] = in[3]; }
  if (J_i) {
    J_i[0] = 1; J_i[1] = 0; J_i[2] = d[0]; J_i[3] = 0;
    J_i[4] = 0; J_i[5] = 1; J_i[6] = 0; J_i[7] = d[1];
  }
}

// IMU axis-alignment intrinsics [c00 c11 c22 c10 c20 c21] -> lower-triangular matrix
// (reference settings.yaml:87-89, inertial.cpp:118-119 asMatrix(), :166 align()).
inline void imu_intrinsics_matrix(const double* c, double* M) {
  M[0] = c[0]; M[1] = 0;    M[2] = 0;
  M[3] = c[3]; M[4] = c[1]; M[5] = 0;
  M[6] = c[4]; M[7] = c[5]; M[8] = c[2];
}
inline void imu_align_jacobian(const double* v, double* J /*3x6*/) {
  std::memset(J, 0, 18 * sizeof(double));
  J[0] = v[0];
  J[6 + 1] = v[1]; J[6 + 3] = v[0];
  J[12 + 2] = v[2]; J[12 + 4] = v[0]; J[12 + 5] = v[1];
}

// ---------------------------------------------------------------------------------------------
// Factor description + layout (reference EvaluatorLayout, evaluators/forward.hpp:19-39).
// ---------------------------------------------------------------------------------------------
enum FactorKind { kPixel = 0, kInertial = 1, kBearing = 2, kManifold = 3 };

// Reference-quirk switches for the inertial evaluator.
enum Quirks {
  kQuirkGyroIntrinsicsInAccelRows = 1,  // (i)  inertial.cpp:136,142,148,158 use I_g where I_a is meant
  kQuirkNoSgJacobian = 2,               // (ii) S_g * a_b_m contributes no state/extrinsics/X_a/gravity Jacobian
  kQuirkLocalExtrinsics = 4,            // (iv) inertial.cpp:157-158 omit R_sb (right-perturbation form)
  kQuirkAxesOffsetsIgnored = 8,         // (v)  inertial.cpp:142,148 use t_bs only (no X_a columns)
  kQuirkAll = 15
};

constexpr int kMaxBlocks = 32;

struct Layout {
  int num_blocks = 0;
  int num_parameters = 0;
  int state_idx = 0, sensor_static_idx = 0, sensor_dynamic_idx = 0, observation_idx = 0;
  int offsets[kMaxBlocks];
  int sizes[kMaxBlocks];
};

struct Factor {
  int kind = kPixel;
  double stamp = 0;
  double measurement[7] = {0, 0, 0, 0, 0, 0, 0};
  int k = 4;     // state spline order (interpolator()->layout().outer.size)
  int k_bg = 4;  // gyroscope bias spline order
  int k_ba = 4;  // accelerometer bias spline order
};

inline int num_residuals(const Factor& f) { return f.kind == kPixel ? 2 : (f.kind == kBearing ? 1 : 6); }
inline int measurement_size(const Factor& f) { return f.kind == kPixel ? 2 : (f.kind == kInertial ? 6 : (f.kind == kBearing ? 3 : 7)); }

// ExteroceptiveCost::update (reference exteroceptive.cpp:25-99): block order
// state (k x 8) | sensor static | sensor dynamic | observation; offsets = exclusive prefix sums.
inline void layout_update(const Factor& f, Layout* L) {
  int n = 0;
  L->state_idx = 0;
  for (int i = 0; i < f.k; ++i) L->sizes[n++] = 8;
  L->sensor_static_idx = n;
  if (f.kind == kManifold) {
    L->sizes[n++] = 7;  // T_bs of a plain Sensor (Traits<Sensor>::kTransformationOffset, reference manifold.cpp:30)
    L->sensor_dynamic_idx = n;
    L->observation_idx = n;  // ManifoldObservation carries no extra variables
  } else if (f.kind == kPixel || f.kind == kBearing) {
    L->sizes[n++] = 7;  // T_bs      Traits<Camera>::kTransformationOffset
    L->sizes[n++] = 4;  // intrinsics
    L->sizes[n++] = 4;  // distortion
    L->sensor_dynamic_idx = n;
    L->observation_idx = n;
    L->sizes[n++] = 3;  // landmark
  } else {
    L->sizes[n++] = 7;  // T_bs
    L->sizes[n++] = 6;  // gyroscope intrinsics
    L->sizes[n++] = 6;  // accelerometer intrinsics
    L->sizes[n++] = 9;  // gyroscope sensitivity S_g
    L->sizes[n++] = 9;  // accelerometer axes offsets X_a
    L->sensor_dynamic_idx = n;
    for (int i = 0; i < f.k_bg; ++i) L->sizes[n++] = 4;
    for (int i = 0; i < f.k_ba; ++i) L->sizes[n++] = 4;
    L->observation_idx = n;
    L->sizes[n++] = 3;  // gravity
  }
  L->num_blocks = n;
  int off = 0;
  for (int i = 0; i < n; ++i) { L->offsets[i] = off; off += L->sizes[i]; }
  L->num_parameters = off;
}

// J_e helper: dst(rows x ncols at column offset) = src (rows x ncols row-major)
inline void put_block(double* J, int ld, int row0, int col0, const double* src, int rows, int cols) {
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) J[(row0 + r) * ld + col0 + c] = src[r * cols + c];
}

// ---------------------------------------------------------------------------------------------
// VisualPixelEvaluator<SE3>::evaluate (reference pixel.cpp:16-146).
// p_ps: parameter blocks; J_e: 2 x num_parameters row-major (zeroed here) or null; p_js: null-flags.
// ---------------------------------------------------------------------------------------------
inline void pixel_evaluate(const Factor& f, const Layout& L, const Basis& basis, const double* const* p_ps, double* J_e,
                           const double* const* p_js, double* prediction) {
  const int o_T_bs = L.sensor_static_idx + 0, o_i = L.sensor_static_idx + 1, o_d = L.sensor_static_idx + 2;
  const int o_p_w = L.observation_idx + 0;
  const double* T_bs = p_ps[o_T_bs];
  const double* c_i = p_ps[o_i];
  const double* c_d = p_ps[o_d];
  const double* p_w = p_ps[o_p_w];
  StateResult S;
  if (!J_e) {
    state_evaluate(basis, p_ps + L.state_idx, f.stamp, 0, false, &S);
    double T_ws[7], T_sw[7], p_s[3], n[2], dn[2];
    se3_group_plus(S.value, T_bs, T_ws, nullptr, nullptr);
    se3_group_inverse(T_ws, T_sw, nullptr);
    se3_vector_plus(T_sw, p_w, p_s, nullptr);
    project_to_plane(p_s, n, nullptr);
    radtan_distort(c_d, n, dn, nullptr, nullptr);
    denormalize(c_i, dn, prediction, nullptr, nullptr);
    return;
  }
  const int np = L.num_parameters;
  std::memset(J_e, 0, sizeof(double) * 2 * np);
  bool J_S_wb = false;
  for (int i = 0; i < f.k; ++i) J_S_wb = J_S_wb || (p_js[L.state_idx + i] != nullptr);
  state_evaluate(basis, p_ps + L.state_idx, f.stamp, 0, J_S_wb, &S);

  double T_ws[7], J_T_ws_S_wb[36], J_T_ws_T_bs[36];
  se3_group_plus(S.value, T_bs, T_ws, J_T_ws_S_wb, J_T_ws_T_bs);
  double T_sw[7], J_T_sw_T_ws[36];
  se3_group_inverse(T_ws, T_sw, J_T_sw_T_ws);
  double p_s[3], J_p_s_T_sw[18];
  se3_vector_plus(T_sw, p_w, p_s, J_p_s_T_sw);
  double n[2], J_n_p[6];
  project_to_plane(p_s, n, J_n_p);

  double dn[2], J_dn_n[4], J_dn_d[8], J_r_dn[4], J_px_i[8];
  radtan_distort(c_d, n, dn, J_dn_n, p_js[o_d] ? J_dn_d : nullptr);
  denormalize(c_i, dn, prediction, J_r_dn, p_js[o_i] ? J_px_i : nullptr);
  if (p_js[o_d]) {
    double t[8];
    mat_mul(J_r_dn, J_dn_d, t, 2, 2, 4);
    put_block(J_e, np, 0, L.offsets[o_d], t, 2, 4);
  }
  if (p_js[o_i]) put_block(J_e, np, 0, L.offsets[o_i], J_px_i, 2, 4);
  double J_r_n[4];
  mat_mul(J_r_dn, J_dn_n, J_r_n, 2, 2, 2);

  double J_r_p_s[6], t26[12], J_r_T_ws[12];
  mat_mul(J_r_n, J_n_p, J_r_p_s, 2, 2, 3);
  mat_mul(J_r_p_s, J_p_s_T_sw, t26, 2, 3, 6);
  mat_mul(t26, J_T_sw_T_ws, J_r_T_ws, 2, 6, 6);

  if (J_S_wb) {
    double t[12];
    mat_mul(J_r_T_ws, J_T_ws_S_wb, t, 2, 6, 6);
    const int cols = 8 * f.k;
    std::vector<double> blk(2 * cols);
    mat_mul(t, S.J[0], blk.data(), 2, 6, cols);
    put_block(J_e, np, 0, L.offsets[L.state_idx], blk.data(), 2, cols);
  }
  if (p_js[o_T_bs]) {
    double t[12], Ad[42], blk[14];
    mat_mul(J_r_T_ws, J_T_ws_T_bs, t, 2, 6, 6);
    se3_adapter(T_bs, Ad);
    mat_mul(t, Ad, blk, 2, 6, 7);
    put_block(J_e, np, 0, L.offsets[o_T_bs], blk, 2, 7);
  }
  if (p_js[o_p_w]) {
    double R_sw[9], blk[6];
    quat_to_rot(T_sw, R_sw);
    mat_mul(J_r_p_s, R_sw, blk, 2, 3, 3);
    put_block(J_e, np, 0, L.offsets[o_p_w], blk, 2, 3);
  }
}

// ---------------------------------------------------------------------------------------------
// InertialEvaluator<SE3>::evaluate (reference inertial.cpp:13-205).  quirks = 0 is the
// mathematically consistent model; quirks = kQuirkAll reproduces the in-tree formulas verbatim.
// The two agree whenever I_g = I_a, S_g = 0, X_a = 0 and R_bs = I -- true for every fixture the
// reference ships (tests/include/tests/sensors/imu.hpp:23-24 except R_bs; settings.yaml:83-106).
// ---------------------------------------------------------------------------------------------
inline void inertial_evaluate(const Factor& f, const Layout& L, const Basis& basis, const Basis& bias_basis,
                              const double* const* p_ps, double* J_e, const double* const* p_js, double* prediction,
                              int quirks) {
  const int o_T_bs = L.sensor_static_idx + 0, o_i_g = L.sensor_static_idx + 1, o_i_a = L.sensor_static_idx + 2;
  const int o_S_g = L.sensor_static_idx + 3, o_X_a = L.sensor_static_idx + 4;
  const int o_b_g = L.sensor_dynamic_idx, o_b_a = o_b_g + f.k_bg, o_g_w = L.observation_idx;
  const double* T_bs = p_ps[o_T_bs];
  const double* S_g_raw = p_ps[o_S_g];  // column-major 3x3 (Eigen::Map default, inertial.cpp:48)
  const double* X_a_raw = p_ps[o_X_a];  // column-major 3x3
  const double* g_w = p_ps[o_g_w];
  double I_g[9], I_a[9], S_g[9];
  imu_intrinsics_matrix(p_ps[o_i_g], I_g);
  imu_intrinsics_matrix(p_ps[o_i_a], I_a);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) S_g[3 * i + j] = S_g_raw[i + 3 * j];

  const bool want_J = (J_e != nullptr);
  bool J_S_wb = false, J_b_g = false, J_b_a = false;
  if (want_J) {
    for (int i = 0; i < f.k; ++i) J_S_wb = J_S_wb || (p_js[L.state_idx + i] != nullptr);
    for (int i = 0; i < f.k_bg; ++i) J_b_g = J_b_g || (p_js[o_b_g + i] != nullptr);
    for (int i = 0; i < f.k_ba; ++i) J_b_a = J_b_a || (p_js[o_b_a + i] != nullptr);
  }
  StateResult S;
  state_evaluate(basis, p_ps + L.state_idx, f.stamp, 2, J_S_wb, &S);
  double b_g[3], b_a[3];
  std::vector<double> J_bg(3 * 4 * f.k_bg), J_ba(3 * 4 * f.k_ba);
  bias_evaluate(bias_basis, p_ps + o_b_g, f.stamp, b_g, J_b_g ? J_bg.data() : nullptr, nullptr);
  bias_evaluate(bias_basis, p_ps + o_b_a, f.stamp, b_a, J_b_a ? J_ba.data() : nullptr, nullptr);

  double R_wb[9], R_bw[9], R_bs[9], R_sb[9];
  quat_to_rot(S.value, R_wb); m3_transpose(R_wb, R_bw);
  quat_to_rot(T_bs, R_bs); m3_transpose(R_bs, R_sb);
  const double* t_bs = T_bs + 4;
  const double* w_b = S.velocity;
  const double* al_b = S.acceleration;
  const double* A_lin = S.acceleration + 3;
  double w_x[9], F_a[9], al_x[9];
  hat(w_b, w_x); hat(al_b, al_x);
  m3_mul(w_x, w_x, F_a); m3_add(F_a, al_x, F_a);
  double Rg[3], a_b_i[3], a_b_m[3];
  m3_vec(R_bw, g_w, Rg);
  for (int c = 0; c < 3; ++c) a_b_i[c] = A_lin[c] - Rg[c];
  double cr[3][3];  // per-axis lever arms c_r = X_a[:, r] + t_bs
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) cr[r][c] = X_a_raw[c + 3 * r] + t_bs[c];
    a_b_m[r] = a_b_i[r] + F_a[3 * r] * cr[r][0] + F_a[3 * r + 1] * cr[r][1] + F_a[3 * r + 2] * cr[r][2];
  }
  double I_g_R_sb[9], I_a_R_sb[9], w_s[3], a_s[3];
  m3_mul(I_g, R_sb, I_g_R_sb);
  m3_mul(I_a, R_sb, I_a_R_sb);
  m3_vec(R_sb, w_b, w_s);
  m3_vec(R_sb, a_b_m, a_s);
  {
    double t1[3], t2[3], t3[3];
    m3_vec(I_g_R_sb, w_b, t1);
    m3_vec(S_g, a_b_m, t2);
    m3_vec(I_a_R_sb, a_b_m, t3);
    for (int c = 0; c < 3; ++c) {
      prediction[c] = t1[c] + t2[c] + b_g[c];
      prediction[3 + c] = t3[c] + b_a[c];
    }
  }
  if (!want_J) return;

  const int np = L.num_parameters;
  std::memset(J_e, 0, sizeof(double) * 6 * np);
  const bool q_i = quirks & kQuirkGyroIntrinsicsInAccelRows;
  const bool q_sg = quirks & kQuirkNoSgJacobian;
  const bool q_loc = quirks & kQuirkLocalExtrinsics;
  const bool q_xa = quirks & kQuirkAxesOffsetsIgnored;
  const double* I_lin_R_sb = q_i ? I_g_R_sb : I_a_R_sb;  // quirk (i)
  const double* I_lin = q_i ? I_g : I_a;

  // d a_b_m / d omega (K_w) and / d alpha (K_al): row r from lever arm c_r (t_bs only under quirk v).
  double K_w[9], K_al[9];
  for (int r = 0; r < 3; ++r) {
    const double* c = q_xa ? t_bs : cr[r];
    double cx[9], t1[9], t2[9];
    hat(c, cx);
    m3_mul(w_x, cx, t1);
    m3_mul(cx, w_x, t2);
    for (int j = 0; j < 3; ++j) {
      K_w[3 * r + j] = -(2.0 * t1[3 * r + j] - t2[3 * r + j]);
      K_al[3 * r + j] = -cx[3 * r + j];
    }
  }

  if (J_S_wb) {
    double J_value[36], J_velocity[36], J_acceleration[36];
    std::memset(J_value, 0, sizeof(J_value));
    std::memset(J_velocity, 0, sizeof(J_velocity));
    std::memset(J_acceleration, 0, sizeof(J_acceleration));
    double ax[9], axR[9], t[9];
    hat(a_b_i, ax);
    m3_mul(ax, R_bw, axR);
    m3_mul(I_lin_R_sb, axR, t);  // inertial.cpp:136
    put_block(J_value, 6, 3, 0, t, 3, 3);
    put_block(J_velocity, 6, 0, 0, I_g_R_sb, 3, 3);  // :141
    m3_mul(I_lin_R_sb, K_w, t);                      // :142
    put_block(J_velocity, 6, 3, 0, t, 3, 3);
    m3_mul(I_lin_R_sb, K_al, t);  // :148
    put_block(J_acceleration, 6, 3, 0, t, 3, 3);
    put_block(J_acceleration, 6, 3, 3, I_a_R_sb, 3, 3);  // :150
    if (!q_sg) {
      // S_g * a_b_m term of r_omega (absent in the reference's Jacobians, inertial.cpp:135,147).
      double s[9];
      m3_mul(S_g, axR, s);
      put_block(J_value, 6, 0, 0, s, 3, 3);
      m3_mul(S_g, K_w, s);
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) J_velocity[6 * i + j] += s[3 * i + j];
      m3_mul(S_g, K_al, s);
      put_block(J_acceleration, 6, 0, 0, s, 3, 3);
      put_block(J_acceleration, 6, 0, 3, S_g, 3, 3);
    }
    const int cols = 8 * f.k;
    std::vector<double> blk(6 * cols), tmp(6 * cols);
    mat_mul(J_value, S.J[0], blk.data(), 6, 6, cols);  // :152
    mat_mul(J_velocity, S.J[1], tmp.data(), 6, 6, cols);
    for (int i = 0; i < 6 * cols; ++i) blk[i] += tmp[i];
    mat_mul(J_acceleration, S.J[2], tmp.data(), 6, 6, cols);
    for (int i = 0; i < 6 * cols; ++i) blk[i] += tmp[i];
    put_block(J_e, np, 0, L.offsets[L.state_idx], blk.data(), 6, cols);
  }

  if (p_js[o_T_bs]) {
    double J_T_bs[36];
    std::memset(J_T_bs, 0, sizeof(J_T_bs));
    double wsx[9], asx[9], t[9];
    hat(w_s, wsx); hat(a_s, asx);
    m3_mul(I_g, wsx, t);  // :157
    if (!q_loc) m3_mul(t, R_sb, t);
    put_block(J_T_bs, 6, 0, 0, t, 3, 3);
    m3_mul(I_lin, asx, t);  // :158
    if (!q_loc) m3_mul(t, R_sb, t);
    put_block(J_T_bs, 6, 3, 0, t, 3, 3);
    m3_mul(I_a_R_sb, F_a, t);  // :160
    put_block(J_T_bs, 6, 3, 3, t, 3, 3);
    if (!q_sg) {
      double s[9], u[9];
      m3_mul(S_g, F_a, s);
      put_block(J_T_bs, 6, 0, 3, s, 3, 3);
      // S_g a_b_m does not depend on R_bs.
      (void)u;
    }
    double Ad[42], blk[42];
    se3_adapter(T_bs, Ad);
    mat_mul(J_T_bs, Ad, blk, 6, 6, 7);
    put_block(J_e, np, 0, L.offsets[o_T_bs], blk, 6, 7);
  }
  if (p_js[o_i_g]) {
    double J[18];
    imu_align_jacobian(w_s, J);
    put_block(J_e, np, 0, L.offsets[o_i_g], J, 3, 6);
  }
  if (p_js[o_i_a]) {
    double J[18];
    imu_align_jacobian(a_s, J);
    put_block(J_e, np, 3, L.offsets[o_i_a], J, 3, 6);
  }
  if (p_js[o_S_g]) {  // :176-187
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) J_e[i * np + L.offsets[o_S_g] + i + 3 * j] = a_b_m[j];
  }
  if (p_js[o_X_a]) {  // :189-194
    for (int r = 0; r < 3; ++r)
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          J_e[(3 + i) * np + L.offsets[o_X_a] + 3 * r + j] = I_a_R_sb[3 * i + r] * F_a[3 * r + j];
          if (!q_sg) J_e[i * np + L.offsets[o_X_a] + 3 * r + j] = S_g[3 * i + r] * F_a[3 * r + j];
        }
  }
  if (J_b_g) put_block(J_e, np, 0, L.offsets[o_b_g], J_bg.data(), 3, 4 * f.k_bg);  // :196
  if (J_b_a) put_block(J_e, np, 3, L.offsets[o_b_a], J_ba.data(), 3, 4 * f.k_ba);  // :197
  if (p_js[o_g_w]) {                                                                  // :198
    double t[9];
    m3_mul(I_a_R_sb, R_bw, t);
    m3_scale(t, -1.0);
    put_block(J_e, np, 3, L.offsets[o_g_w], t, 3, 3);
    if (!q_sg) {
      m3_mul(S_g, R_bw, t);
      m3_scale(t, -1.0);
      put_block(J_e, np, 0, L.offsets[o_g_w], t, 3, 3);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// VisualBearingEvaluator<SE3>::evaluate (reference internal/hyper/optimizers/evaluators/bearing.cpp:14-79):
// prediction = landmark position in the sensor frame p_s (3); J_e 3 x num_parameters.
// ---------------------------------------------------------------------------------------------
inline void bearing_evaluate(const Factor& f, const Layout& L, const Basis& basis, const double* const* p_ps, double* J_e,
                             const double* const* p_js, double* prediction) {
  const int o_T_bs = L.sensor_static_idx + 0, o_p_w = L.observation_idx + 0;
  const double* T_bs = p_ps[o_T_bs];
  const double* p_w = p_ps[o_p_w];
  StateResult S;
  if (!J_e) {
    state_evaluate(basis, p_ps + L.state_idx, f.stamp, 0, false, &S);
    double T_ws[7], T_sw[7];
    se3_group_plus(S.value, T_bs, T_ws, nullptr, nullptr);
    se3_group_inverse(T_ws, T_sw, nullptr);
    se3_vector_plus(T_sw, p_w, prediction, nullptr);
    return;
  }
  const int np = L.num_parameters;
  std::memset(J_e, 0, sizeof(double) * 3 * np);
  bool J_S_wb = false;
  for (int i = 0; i < f.k; ++i) J_S_wb = J_S_wb || (p_js[L.state_idx + i] != nullptr);
  state_evaluate(basis, p_ps + L.state_idx, f.stamp, 0, J_S_wb, &S);
  double T_ws[7], J_T_ws_S_wb[36], J_T_ws_T_bs[36], T_sw[7], J_T_sw_T_ws[36], J_p_s_T_sw[18];
  se3_group_plus(S.value, T_bs, T_ws, J_T_ws_S_wb, J_T_ws_T_bs);
  se3_group_inverse(T_ws, T_sw, J_T_sw_T_ws);
  se3_vector_plus(T_sw, p_w, prediction, J_p_s_T_sw);
  double J_p_s_T_ws[18];
  mat_mul(J_p_s_T_sw, J_T_sw_T_ws, J_p_s_T_ws, 3, 6, 6);  // bearing.cpp:71
  if (J_S_wb) {                                            // :73
    double t[18];
    mat_mul(J_p_s_T_ws, J_T_ws_S_wb, t, 3, 6, 6);
    const int cols = 8 * f.k;
    std::vector<double> blk(3 * cols);
    mat_mul(t, S.J[0], blk.data(), 3, 6, cols);
    put_block(J_e, np, 0, L.offsets[L.state_idx], blk.data(), 3, cols);
  }
  if (p_js[o_T_bs]) {  // :74
    double t[18], Ad[42], blk[21];
    mat_mul(J_p_s_T_ws, J_T_ws_T_bs, t, 3, 6, 6);
    se3_adapter(T_bs, Ad);
    mat_mul(t, Ad, blk, 3, 6, 7);
    put_block(J_e, np, 0, L.offsets[o_T_bs], blk, 3, 7);
  }
  if (p_js[o_p_w]) {  // :75
    double R_sw[9];
    quat_to_rot(T_sw, R_sw);
    put_block(J_e, np, 0, L.offsets[o_p_w], R_sw, 3, 3);
  }
}

// ---------------------------------------------------------------------------------------------
// ManifoldEvaluator<SE3>::evaluate (reference internal/hyper/optimizers/evaluators/manifold.cpp:12-61):
// prediction = T_ws = T_wb (+) T_bs (7); J_e 6 x num_parameters (rows = tangent of the prediction).
// ---------------------------------------------------------------------------------------------
inline void manifold_evaluate(const Factor& f, const Layout& L, const Basis& basis, const double* const* p_ps, double* J_e,
                              const double* const* p_js, double* prediction) {
  const int o_T_bs = L.sensor_static_idx + 0;
  const double* T_bs = p_ps[o_T_bs];
  StateResult S;
  if (!J_e) {
    state_evaluate(basis, p_ps + L.state_idx, f.stamp, 0, false, &S);
    se3_group_plus(S.value, T_bs, prediction, nullptr, nullptr);
    return;
  }
  const int np = L.num_parameters;
  std::memset(J_e, 0, sizeof(double) * 6 * np);
  bool J_S_wb = false;
  for (int i = 0; i < f.k; ++i) J_S_wb = J_S_wb || (p_js[L.state_idx + i] != nullptr);
  state_evaluate(basis, p_ps + L.state_idx, f.stamp, 0, J_S_wb, &S);
  double J_r_S_wb[36], J_r_T_bs[36];
  se3_group_plus(S.value, T_bs, prediction, J_r_S_wb, J_r_T_bs);
  if (J_S_wb) {  // manifold.cpp:56
    const int cols = 8 * f.k;
    std::vector<double> blk(6 * cols);
    mat_mul(J_r_S_wb, S.J[0], blk.data(), 6, 6, cols);
    put_block(J_e, np, 0, L.offsets[L.state_idx], blk.data(), 6, cols);
  }
  if (p_js[o_T_bs]) {  // :57
    double Ad[42], blk[42];
    se3_adapter(T_bs, Ad);
    mat_mul(J_r_T_bs, Ad, blk, 6, 6, 7);
    put_block(J_e, np, 0, L.offsets[o_T_bs], blk, 6, 7);
  }
}

// ---------------------------------------------------------------------------------------------
// Metrics (HyperVariables, not in the tree -- [INFERRED]):
//   AngularMetric<Bearing>::distance(lhs, rhs, J_lhs): the angle between the two vectors,
//     theta = atan2(|lhs x rhs|, lhs . rhs)  (1 residual; wired at reference optimizer.cpp:192),
//     J_lhs = [c (rhs x n) - s rhs]^T / (|lhs|^2 |rhs|^2), n = (lhs x rhs)/s; zero when s -> 0.
//   ManifoldMetric<SE3>::distance(lhs, rhs, J_lhs): lhs (-) rhs in the tangent convention of this
//     repo, [Log(R_lhs R_rhs^T) | p_lhs - p_rhs]  (6 residuals; wired at optimizer.cpp:237),
//     J_lhs = blockdiag(Jl^{-1}(theta), I) with Jl^{-1}(theta) = Jr^{-1}(-theta).
// ---------------------------------------------------------------------------------------------
inline double angular_distance(const double* a, const double* b, double* J_a /*1x3 or null*/) {
  double cr[3];
  v3_cross(a, b, cr);
  const double s = std::sqrt(v3_dot(cr, cr)), c = v3_dot(a, b);
  const double theta = std::atan2(s, c);
  if (J_a) {
    if (s < 1e-300) { J_a[0] = J_a[1] = J_a[2] = 0.0; }
    else {
      const double n[3] = {cr[0] / s, cr[1] / s, cr[2] / s};
      double bn[3];
      v3_cross(b, n, bn);
      const double den = v3_dot(a, a) * v3_dot(b, b);
      for (int i = 0; i < 3; ++i) J_a[i] = (c * bn[i] - s * b[i]) / den;
    }
  }
  return theta;
}
inline void manifold_distance(const double* lhs, const double* rhs, double* d /*6*/, double* J_lhs /*6x6 or null*/) {
  double qc[4], qr[4];
  quat_conj(rhs, qc);
  quat_mul(lhs, qc, qr);   // R_lhs R_rhs^T
  quat_log(qr, d);
  for (int i = 0; i < 3; ++i) d[3 + i] = lhs[4 + i] - rhs[4 + i];
  if (J_lhs) {
    std::memset(J_lhs, 0, 36 * sizeof(double));
    const double neg[3] = {-d[0], -d[1], -d[2]};
    double Jli[9];
    so3_Jr_inv(neg, Jli);
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) J_lhs[6 * i + j] = Jli[3 * i + j];
      J_lhs[6 * (3 + i) + 3 + i] = 1.0;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// ExteroceptiveCost<CERES>::Evaluate (reference exteroceptive.cpp:101-160) with
// CartesianMetric (distance = lhs - rhs, J_lhs = I; wired at optimizer.cpp:215,256) and no weights.
// jacobians[i] row-major num_residuals x sizes[i], ambient coordinates; may be null.
// ---------------------------------------------------------------------------------------------
inline bool cost_evaluate(const Factor& f, const Layout& L, const Basis& basis, const Basis& bias_basis,
                          const double* const* parameters, double* residuals, double** jacobians, int quirks) {
  const int nr = num_residuals(f);
  const int ne = (f.kind == kPixel) ? 2 : (f.kind == kBearing ? 3 : 6);   // rows of the evaluator Jacobian J_e
  double prediction[7];
  std::vector<double> J_e;
  double* pJ = nullptr;
  if (jacobians) { J_e.resize((size_t)ne * L.num_parameters); pJ = J_e.data(); }
  switch (f.kind) {
    case kPixel: pixel_evaluate(f, L, basis, parameters, pJ, jacobians, prediction); break;
    case kInertial: inertial_evaluate(f, L, basis, bias_basis, parameters, pJ, jacobians, prediction, quirks); break;
    case kBearing: bearing_evaluate(f, L, basis, parameters, pJ, jacobians, prediction); break;
    default: manifold_evaluate(f, L, basis, parameters, pJ, jacobians, prediction); break;
  }
  // metric: CartesianMetric (pixel, inertial: optimizer.cpp:215,256), AngularMetric (bearing, :192),
  // ManifoldMetric (manifold, :237); J_w = J_m * J_e (exteroceptive.cpp:136-137)
  double J_m[36];
  if (f.kind == kPixel || f.kind == kInertial) {
    for (int i = 0; i < nr; ++i) residuals[i] = prediction[i] - f.measurement[i];
  } else if (f.kind == kBearing) {
    residuals[0] = angular_distance(prediction, f.measurement, jacobians ? J_m : nullptr);
  } else {
    manifold_distance(prediction, f.measurement, residuals, jacobians ? J_m : nullptr);
  }
  if (!jacobians) return true;
  std::vector<double> J_w;
  const double* Jw = J_e.data();
  if (f.kind == kBearing || f.kind == kManifold) {
    J_w.resize((size_t)nr * L.num_parameters);
    mat_mul(J_m, J_e.data(), J_w.data(), nr, ne, L.num_parameters);
    Jw = J_w.data();
  }
  for (int b = 0; b < L.num_blocks; ++b) {
    if (!jacobians[b]) continue;
    const int sz = L.sizes[b];
    for (int r = 0; r < nr; ++r)
      for (int c = 0; c < sz; ++c) jacobians[b][r * sz + c] = Jw[(size_t)r * L.num_parameters + L.offsets[b] + c];
  }
  return true;
}

}  // namespace ho
