// TEST INFRASTRUCTURE ONLY (see ho_math.h header).  PARITY UNPINNED (HyperState is not in the tree).
//
// Restates what the reference calls through
//   AbstractState::evaluate(StateQuery{stamp, derivative, jacobian}, const Scalar* const*)
//   (call sites: reference internal/hyper/optimizers/evaluators/pixel.cpp:55-56,74-75,
//    inertial.cpp:54-55,93-94) with BasisInterpolator(degree, uniform=true)
//   (reference tests/internal/tests/optimizers/evaluators/pixel.cpp:50) and
//   ManifoldPolicy<Stamped<SE3>> (reference internal/hyper/optimizers/abstract.cpp:79-80).
//
// Uniform cumulative B-spline of order k (degree k-1) on SO(3) x R^3 (Sommer et al. 2020, cited
// reference README.md:174):
//   R(t) = R_0 * prod_{j=1}^{k-1} Exp(lambda_j(u) d_j),  d_j = Log(R_{j-1}^T R_j)
//   p(t) = p_0 + sum_j lambda_j(u) (p_j - p_{j-1})
// Control point m of a query is parameter block m: [qx qy qz qw px py pz | stamp] (8 doubles).
// The segment is recovered from the blocks' own stamps: with left = (k-1)/2 (reference
// abstract.cpp:89 places point i at lower + (i - (k-1)/2) * separation) the query stamp lies in
// [stamp(left), stamp(left+1)), u = (t - stamp(left)) / (stamp(left+1) - stamp(left)).
//
// Outputs mirror the reference's StateResult: derivatives[0] = pose (7), [1] = velocity tangent
// [omega_b | R^T pdot], [2] = acceleration tangent [alpha_b | R^T pddot]; jacobians[d] = 6 x 8k
// (row-major here) in AMBIENT coordinates of the k blocks, stamp columns zero.
// Linear rows of jacobians[1], jacobians[2] are taken "at fixed R" (translation columns only):
// this is what reference inertial.cpp:125,136 requires for J_value = I R_sb hat(a_b_i) R_bw to be
// the correct chain-rule partial (SURVEY.md section 8a quirk iii).
#pragma once
#include "ho_math.h"

namespace ho {

constexpr int kMaxOrder = 6;

struct Basis {
  int k = 0;
  double Mc[kMaxOrder][kMaxOrder];  // cumulative blending matrix: lambda_j(u) = sum_n Mc[j][n] u^n
};

inline double binom(int n, int r) {
  if (r < 0 || r > n) return 0;
  double v = 1;
  for (int i = 1; i <= r; ++i) v = v * (n - r + i) / i;
  return v;
}

// Blending matrix of the uniform B-spline of order k and its cumulative form.
inline void basis_init(Basis* b, int k) {
  b->k = k;
  double M[kMaxOrder][kMaxOrder];
  double fact = 1;
  for (int i = 2; i <= k - 1; ++i) fact *= i;
  for (int s = 0; s < k; ++s)
    for (int n = 0; n < k; ++n) {
      double sum = 0;
      for (int l = s; l <= k - 1; ++l) {
        const int e = k - 1 - n;
        const double base = (double)(k - 1 - l);
        double pw = 1;
        for (int i = 0; i < e; ++i) pw *= base;  // 0^0 = 1
        sum += ((l - s) % 2 ? -1.0 : 1.0) * binom(k, l - s) * pw;
      }
      M[s][n] = binom(k - 1, n) / fact * sum;
    }
  for (int j = 0; j < k; ++j)
    for (int n = 0; n < k; ++n) {
      double s = 0;
      for (int r = j; r < k; ++r) s += M[r][n];
      b->Mc[j][n] = s;
    }
}

// lam[d][j], d = 0 (value), 1 (d/dt), 2 (d2/dt2); j = 0..k-1 (lam[0][0] = 1).  lam[d][k] = 0.
inline void basis_eval(const Basis& b, double u, double inv_dt, double lam[3][kMaxOrder + 1]) {
  const int k = b.k;
  for (int j = 0; j < k; ++j) {
    double v = 0, d1 = 0, d2 = 0;
    for (int n = k - 1; n >= 0; --n) {  // Horner
      d2 = d2 * u + 2.0 * d1;
      d1 = d1 * u + v;
      v = v * u + b.Mc[j][n];
    }
    lam[0][j] = v;
    lam[1][j] = d1 * inv_dt;
    lam[2][j] = d2 * inv_dt * inv_dt;
  }
  lam[0][0] = 1.0; lam[1][0] = 0.0; lam[2][0] = 0.0;  // exact
  lam[0][k] = lam[1][k] = lam[2][k] = 0.0;
}

// d theta_global / d q (3x4): theta = 2 vec(dq (x) q^*) = 2 [w I + v^ | -v] dq.
// This is the rotation part of the reference's SE3JacobianAdapter / SU2JacobianAdapter
// (reference pixel.cpp:141, manifolds/variables/se3.cpp:15-17): J_tangent * Adapter is an
// ambient Jacobian whose product with Ceres' EigenQuaternionManifold PlusJacobian is dr/d(delta).
inline void su2_adapter(const double* q, double* A /*3x4*/) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  A[0] = 2 * w;  A[1] = -2 * z; A[2] = 2 * y;  A[3] = -2 * x;
  A[4] = 2 * z;  A[5] = 2 * w;  A[6] = -2 * x; A[7] = -2 * y;
  A[8] = -2 * y; A[9] = 2 * x;  A[10] = 2 * w; A[11] = -2 * z;
}

struct StateResult {
  double value[7];
  double velocity[6];
  double acceleration[6];
  // jacobians[d]: 6 x (8k) row-major, leading dimension 8k.
  double J[3][6 * 8 * kMaxOrder];
  // Tangent-space by-products (used by the compact/batched restatement and by tests):
  double R[9];
};

// derivative in {0, 2}; jac = whether Jacobians are requested.
inline void state_evaluate(const Basis& basis, const double* const* cps, double stamp, int derivative, bool jac,
                           StateResult* out) {
  const int k = basis.k;
  const int left = (k - 1) / 2;
  const double t0 = cps[left][7], t1 = cps[left + 1][7];
  const double inv_dt = 1.0 / (t1 - t0);
  const double u = (stamp - t0) * inv_dt;
  double lam[3][kMaxOrder + 1];
  basis_eval(basis, u, inv_dt, lam);

  double R[kMaxOrder][9], d[kMaxOrder][3], G[kMaxOrder][9];
  for (int m = 0; m < k; ++m) quat_to_rot(cps[m], R[m]);
  for (int j = 1; j < k; ++j) {
    double qc[4], qr[4], Ji[9];
    quat_conj(cps[j - 1], qc);
    quat_mul(qc, cps[j], qr);
    quat_log(qr, d[j]);
    so3_Jr_inv(d[j], Ji);
    m3_mult(Ji, R[j], G[j]);  // G_j = Jr^{-1}(d_j) R_j^T
  }

  // Value.
  double q[4] = {cps[0][0], cps[0][1], cps[0][2], cps[0][3]};
  double A[kMaxOrder][9], P[kMaxOrder][9];
  m3_copy(R[0], P[0]);
  for (int j = 1; j < k; ++j) {
    double w[3] = {lam[0][j] * d[j][0], lam[0][j] * d[j][1], lam[0][j] * d[j][2]};
    double qe[4];
    quat_exp(w, qe);
    quat_mul(q, qe, q);
    so3_exp(w, A[j]);
    m3_mul(P[j - 1], A[j], P[j]);
  }
  double p[3] = {cps[0][4], cps[0][5], cps[0][6]}, pd[3] = {0, 0, 0}, pdd[3] = {0, 0, 0};
  for (int j = 1; j < k; ++j)
    for (int c = 0; c < 3; ++c) {
      const double dp = cps[j][4 + c] - cps[j - 1][4 + c];
      p[c] += lam[0][j] * dp;
      pd[c] += lam[1][j] * dp;
      pdd[c] += lam[2][j] * dp;
    }
  for (int c = 0; c < 4; ++c) out->value[c] = q[c];
  for (int c = 0; c < 3; ++c) out->value[4 + c] = p[c];
  const double* Rt = P[k - 1];
  m3_copy(Rt, out->R);

  // Body rates (Sommer et al. recursion, right-multiplied increments).
  double w_at[kMaxOrder][3], ATw[kMaxOrder][3], ATwd[kMaxOrder][3];
  double w[3] = {0, 0, 0}, wd[3] = {0, 0, 0};
  if (derivative >= 1) {
    for (int j = 1; j < k; ++j) {
      m3_tvec(A[j], w, ATw[j]);
      m3_tvec(A[j], wd, ATwd[j]);
      double wn[3], cr[3];
      for (int c = 0; c < 3; ++c) wn[c] = ATw[j][c] + lam[1][j] * d[j][c];
      v3_cross(wn, d[j], cr);
      for (int c = 0; c < 3; ++c) wd[c] = ATwd[j][c] + lam[1][j] * cr[c] + lam[2][j] * d[j][c];
      v3_copy(wn, w);
      v3_copy(wn, w_at[j]);
    }
    double v[3], a[3];
    m3_tvec(Rt, pd, v);
    m3_tvec(Rt, pdd, a);
    for (int c = 0; c < 3; ++c) {
      out->velocity[c] = w[c]; out->velocity[3 + c] = v[c];
      out->acceleration[c] = wd[c]; out->acceleration[3 + c] = a[c];
    }
  }
  if (!jac) return;

  const int ld = 8 * k;
  const int nd = (derivative >= 2) ? 3 : 1;
  for (int dd = 0; dd < nd; ++dd) std::memset(out->J[dd], 0, sizeof(double) * 6 * ld);

  // d(.)/d d_j for theta (T), omega (X), alpha (Y).
  double T[kMaxOrder][9], X[kMaxOrder][9], Y[kMaxOrder][9];
  for (int j = 1; j < k; ++j) {
    double wj[3] = {lam[0][j] * d[j][0], lam[0][j] * d[j][1], lam[0][j] * d[j][2]};
    double Jr[9];
    so3_Jr(wj, Jr);
    double lJr[9];
    m3_copy(Jr, lJr); m3_scale(lJr, lam[0][j]);
    m3_mul(P[j], lJr, T[j]);
    if (nd > 1) {
      double H[9], Xj[9], Yj[9], Hd[9], Hw[9], Dh[9], tmp[9];
      hat(ATw[j], H);
      m3_mul(H, lJr, Xj);
      for (int i = 0; i < 3; ++i) Xj[4 * i] += lam[1][j];
      hat(ATwd[j], Hd);
      m3_mul(Hd, lJr, Yj);
      hat(w_at[j], Hw);
      hat(d[j], Dh);
      m3_mul(Dh, Xj, tmp);
      for (int i = 0; i < 9; ++i) Yj[i] += lam[1][j] * (Hw[i] - tmp[i]);
      for (int i = 0; i < 3; ++i) Yj[4 * i] += lam[2][j];
      for (int l = j + 1; l < k; ++l) {
        double Xn[9], Yn[9], Dl[9], t2[9];
        m3_tmul(A[l], Xj, Xn);
        m3_tmul(A[l], Yj, Yn);
        hat(d[l], Dl);
        m3_mul(Dl, Xn, t2);
        for (int i = 0; i < 9; ++i) Yn[i] -= lam[1][l] * t2[i];
        m3_copy(Xn, Xj); m3_copy(Yn, Yj);
      }
      m3_copy(Xj, X[j]); m3_copy(Yj, Y[j]);
    }
  }

  double RtT[9];
  m3_transpose(Rt, RtT);
  for (int m = 0; m < k; ++m) {
    double Aq[12];
    su2_adapter(cps[m], Aq);
    for (int dd = 0; dd < nd; ++dd) {
      double (*D)[9] = (dd == 0) ? T : (dd == 1 ? X : Y);
      double Dphi[9];
      m3_zero(Dphi);
      if (m >= 1) { double t[9]; m3_mul(D[m], G[m], t); m3_add(Dphi, t, Dphi); }
      if (m + 1 < k) { double t[9]; m3_mul(D[m + 1], G[m + 1], t); m3_sub(Dphi, t, Dphi); }
      if (dd == 0 && m == 0) for (int i = 0; i < 3; ++i) Dphi[4 * i] += 1.0;
      // rotation columns (ambient): Dphi (3x3) * Aq (3x4)
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c)
          out->J[dd][r * ld + 8 * m + c] = Dphi[3 * r] * Aq[c] + Dphi[3 * r + 1] * Aq[4 + c] + Dphi[3 * r + 2] * Aq[8 + c];
      // translation columns
      const double wgt = lam[dd][m] - lam[dd][m + 1];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
          out->J[dd][(3 + r) * ld + 8 * m + 4 + c] = (dd == 0) ? (r == c ? wgt : 0.0) : wgt * RtT[3 * r + c];
    }
  }
}

// Euclidean (bias) spline: blocks [b(3) | stamp]; value (3) and 3 x 4k ambient Jacobian.
// Restates imu.gyroscopeBias().evaluate(StateQuery{stamp, kValueIndex, J}, ptrs)
// (reference inertial.cpp:59-60,100-108).
inline void bias_evaluate(const Basis& basis, const double* const* cps, double stamp, double* value, double* J /*3 x 4k or null*/,
                          double* weights /*k or null*/) {
  const int k = basis.k;
  const int left = (k - 1) / 2;
  const double t0 = cps[left][3], t1 = cps[left + 1][3];
  const double inv_dt = 1.0 / (t1 - t0);
  const double u = (stamp - t0) * inv_dt;
  double lam[3][kMaxOrder + 1];
  basis_eval(basis, u, inv_dt, lam);
  value[0] = value[1] = value[2] = 0;
  if (J) std::memset(J, 0, sizeof(double) * 3 * 4 * k);
  for (int m = 0; m < k; ++m) {
    const double wgt = lam[0][m] - lam[0][m + 1];
    if (weights) weights[m] = wgt;
    for (int c = 0; c < 3; ++c) {
      value[c] += wgt * cps[m][c];
      if (J) J[c * 4 * k + 4 * m + c] = wgt;
    }
  }
}

}  // namespace ho
