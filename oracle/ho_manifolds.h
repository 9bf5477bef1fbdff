// TEST INFRASTRUCTURE ONLY (see ho_math.h header).  PARITY UNPINNED: Ceres is not installed;
// semantics restated from the Ceres 2.1 public documentation / headers (manifold.h,
// product_manifold.h, sphere_manifold.h) for exactly the manifolds the reference instantiates:
//   SU2        -> ceres::EigenQuaternionManifold or all-constant SubsetManifold
//                 (reference internal/hyper/optimizers/ceres/manifolds/variables/su2.cpp:17-23)
//   SE3        -> ProductManifold<SU2-manifold, R3-manifold>                  (se3.cpp:19-24)
//   Stamped<T> -> ProductManifold<T-manifold, stamp-manifold>                 (stamped.hpp:31-37)
//   Cartesian  -> EuclideanManifold or SubsetManifold(all constant)           (euclidean.hpp:34-40)
//   Gravity/Bearing -> SphereManifold<3>                                      (bearing.cpp:11-17)
// A manifold is a product of parts; the 7 hooks of reference wrapper.hpp:24-50 are provided.
#pragma once
#include <vector>

#include "ho_math.h"

namespace ho {

enum PartKind { kEuclidean = 0, kConstant = 1, kQuaternion = 2, kSphere = 3 };

struct Part { int kind; int ambient; };

struct Manifold {
  std::vector<Part> parts;
  int ambient_size() const { int n = 0; for (auto& p : parts) n += p.ambient; return n; }
  int tangent_size() const {
    int n = 0;
    for (auto& p : parts) n += (p.kind == kEuclidean) ? p.ambient : (p.kind == kConstant ? 0 : p.ambient - 1);
    return n;
  }
};

inline int part_tangent(const Part& p) { return p.kind == kEuclidean ? p.ambient : (p.kind == kConstant ? 0 : p.ambient - 1); }

// Householder vector of ceres/internal/sphere_manifold_functions.h (ComputeHouseholderVector).
inline void householder(const double* x, int n, double* v, double* beta) {
  double sigma = 0;
  for (int i = 0; i < n - 1; ++i) sigma += x[i] * x[i];
  for (int i = 0; i < n; ++i) v[i] = x[i];
  v[n - 1] = 1.0;
  *beta = 0.0;
  const double x_pivot = x[n - 1];
  if (sigma <= 2.220446049250313e-16) {
    if (x_pivot < 0.0) *beta = 2.0;
    return;
  }
  const double mu = std::sqrt(x_pivot * x_pivot + sigma);
  double v_pivot;
  if (x_pivot <= 0.0) v_pivot = x_pivot - mu;
  else v_pivot = -sigma / (x_pivot + mu);
  *beta = 2.0 * v_pivot * v_pivot / (sigma + v_pivot * v_pivot);
  for (int i = 0; i < n - 1; ++i) v[i] /= v_pivot;
}

inline void part_plus(const Part& p, const double* x, const double* delta, double* out) {
  const int n = p.ambient;
  switch (p.kind) {
    case kEuclidean: for (int i = 0; i < n; ++i) out[i] = x[i] + delta[i]; break;
    case kConstant: for (int i = 0; i < n; ++i) out[i] = x[i]; break;
    case kQuaternion: {
      // EigenQuaternionManifold: x_plus = [sin(|d|)/|d| d, cos|d|] (x) x   (storage x y z w)
      const double nd = std::sqrt(v3_dot(delta, delta));
      double q[4];
      if (nd == 0.0) { q[0] = q[1] = q[2] = 0; q[3] = 1; }
      else { const double s = std::sin(nd) / nd; q[0] = s * delta[0]; q[1] = s * delta[1]; q[2] = s * delta[2]; q[3] = std::cos(nd); }
      quat_mul(q, x, out);
    } break;
    case kSphere: {
      double nd = 0;
      for (int i = 0; i < n - 1; ++i) nd += delta[i] * delta[i];
      nd = std::sqrt(nd);
      if (nd == 0.0) { for (int i = 0; i < n; ++i) out[i] = x[i]; break; }
      double v[8], beta, y[8], nx = 0;
      householder(x, n, v, &beta);
      for (int i = 0; i < n; ++i) nx += x[i] * x[i];
      nx = std::sqrt(nx);
      const double s = std::sin(nd) / nd;
      for (int i = 0; i < n - 1; ++i) y[i] = s * delta[i];
      y[n - 1] = std::cos(nd);
      double vy = 0;
      for (int i = 0; i < n; ++i) vy += v[i] * y[i];
      for (int i = 0; i < n; ++i) out[i] = nx * (y[i] - v[i] * beta * vy);
    } break;
  }
}

// J: ambient x tangent row-major with leading dimension ld, written at (row0, col0).
inline void part_plus_jacobian(const Part& p, const double* x, double* J, int ld, int row0, int col0) {
  const int n = p.ambient;
  switch (p.kind) {
    case kEuclidean: for (int i = 0; i < n; ++i) J[(row0 + i) * ld + col0 + i] = 1.0; break;
    case kConstant: break;
    case kQuaternion: {
      const double qx = x[0], qy = x[1], qz = x[2], qw = x[3];
      const double M[12] = {qw, qz, -qy, -qz, qw, qx, qy, -qx, qw, -qx, -qy, -qz};
      for (int i = 0; i < 4; ++i) for (int j = 0; j < 3; ++j) J[(row0 + i) * ld + col0 + j] = M[3 * i + j];
    } break;
    case kSphere: {
      double v[8], beta, nx = 0;
      householder(x, n, v, &beta);
      for (int i = 0; i < n; ++i) nx += x[i] * x[i];
      nx = std::sqrt(nx);
      for (int c = 0; c < n - 1; ++c)
        for (int r = 0; r < n; ++r) J[(row0 + r) * ld + col0 + c] = nx * ((r == c ? 1.0 : 0.0) - beta * v[c] * v[r]);
    } break;
  }
}

inline void part_minus(const Part& p, const double* y, const double* x, double* out) {
  const int n = p.ambient;
  switch (p.kind) {
    case kEuclidean: for (int i = 0; i < n; ++i) out[i] = y[i] - x[i]; break;
    case kConstant: break;
    case kQuaternion: {
      // y (x) x^{-1} = [sin(|d|) d/|d|, cos|d|]
      double xc[4], q[4];
      quat_conj(x, xc);
      quat_mul(y, xc, q);
      const double u = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
      if (u > 0.0) { const double th = std::atan2(u, q[3]); for (int i = 0; i < 3; ++i) out[i] = th * q[i] / u; }
      else { out[0] = out[1] = out[2] = 0; }
    } break;
    case kSphere: {
      double v[8], beta, hy[8], nx = 0;
      householder(x, n, v, &beta);
      for (int i = 0; i < n; ++i) nx += x[i] * x[i];
      nx = std::sqrt(nx);
      double vy = 0;
      for (int i = 0; i < n; ++i) vy += v[i] * y[i];
      for (int i = 0; i < n; ++i) hy[i] = y[i] - v[i] * beta * vy;
      double y_last = hy[n - 1], nh = 0;
      for (int i = 0; i < n - 1; ++i) nh += hy[i] * hy[i];
      nh = std::sqrt(nh);
      if (nh == 0.0) { for (int i = 0; i < n - 1; ++i) out[i] = 0; }
      else { const double th = std::atan2(nh, y_last); for (int i = 0; i < n - 1; ++i) out[i] = th * hy[i] / nh; }
      (void)nx;
    } break;
  }
}

inline void manifold_plus(const Manifold& m, const double* x, const double* delta, double* out) {
  int a = 0, t = 0;
  for (auto& p : m.parts) { part_plus(p, x + a, delta + t, out + a); a += p.ambient; t += part_tangent(p); }
}
// J: ambient x tangent, row-major, zero-filled here.
inline void manifold_plus_jacobian(const Manifold& m, const double* x, double* J) {
  const int na = m.ambient_size(), nt = m.tangent_size();
  for (int i = 0; i < na * nt; ++i) J[i] = 0.0;
  int a = 0, t = 0;
  for (auto& p : m.parts) { part_plus_jacobian(p, x + a, J, nt, a, t); a += p.ambient; t += part_tangent(p); }
}
inline void manifold_minus(const Manifold& m, const double* y, const double* x, double* out) {
  int a = 0, t = 0;
  for (auto& p : m.parts) { part_minus(p, y + a, x + a, out + t); a += p.ambient; t += part_tangent(p); }
}

// Manifold factories mirroring the reference's wrappers.
inline Manifold make_su2(bool constant) { return Manifold{{{constant ? kConstant : kQuaternion, 4}}}; }
inline Manifold make_se3(bool rot_const, bool trans_const) {
  return Manifold{{{rot_const ? kConstant : kQuaternion, 4}, {trans_const ? kConstant : kEuclidean, 3}}};
}
inline Manifold make_stamped_se3(bool time_const, bool rot_const, bool trans_const) {
  Manifold m = make_se3(rot_const, trans_const);
  m.parts.push_back({time_const ? kConstant : kEuclidean, 1});
  return m;
}
inline Manifold make_cartesian(int n, bool constant) { return Manifold{{{constant ? kConstant : kEuclidean, n}}}; }
inline Manifold make_stamped_cartesian(int n, bool time_const, bool constant) {
  return Manifold{{{constant ? kConstant : kEuclidean, n}, {time_const ? kConstant : kEuclidean, 1}}};
}
inline Manifold make_sphere(int n, bool constant) { return Manifold{{{constant ? kConstant : kSphere, n}}}; }

}  // namespace ho
