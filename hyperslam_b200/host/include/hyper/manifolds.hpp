// hyper::Manifold -- the seven hooks of the reference's ManifoldWrapper
// (reference include/hyper/optimizers/ceres/manifolds/variables/wrapper.hpp:24-50) on the manifolds
// the reference instantiates (su2.cpp:17-23, se3.cpp:19-24, stamped.hpp:31-37, euclidean.hpp:34-40,
// bearing.cpp:11-17).  Semantics are Ceres 2.1's (EigenQuaternionManifold, SphereManifold<3>,
// EuclideanManifold, SubsetManifold, ProductManifold).  Row-major matrices.
#pragma once
#include <cmath>
#include <vector>

#include "hyper/variables.hpp"

namespace hyper {

class Manifold {
 public:
  enum Kind { kEuclidean, kConstant, kQuaternion, kSphere };
  struct Part { Kind kind; int ambient; };

  static Manifold SU2(bool constant = false) { return Manifold{{{constant ? kConstant : kQuaternion, 4}}}; }
  static Manifold SE3(bool rotation_constant = false, bool translation_constant = false) {
    return Manifold{{{rotation_constant ? kConstant : kQuaternion, 4}, {translation_constant ? kConstant : kEuclidean, 3}}};
  }
  static Manifold StampedSE3(bool time_constant = true, bool rotation_constant = false, bool translation_constant = false) {
    Manifold m = SE3(rotation_constant, translation_constant);
    m.parts_.push_back({time_constant ? kConstant : kEuclidean, 1});
    return m;
  }
  static Manifold Euclidean(int n, bool constant = false) { return Manifold{{{constant ? kConstant : kEuclidean, n}}}; }
  static Manifold StampedEuclidean(int n, bool time_constant = true, bool constant = false) {
    return Manifold{{{constant ? kConstant : kEuclidean, n}, {time_constant ? kConstant : kEuclidean, 1}}};
  }
  static Manifold Sphere(int n = 3, bool constant = false) { return Manifold{{{constant ? kConstant : kSphere, n}}}; }

  int AmbientSize() const { int n = 0; for (auto& p : parts_) n += p.ambient; return n; }
  int TangentSize() const { int n = 0; for (auto& p : parts_) n += tangent(p); return n; }
  bool Plus(const Scalar* x, const Scalar* delta, Scalar* x_plus_delta) const;
  bool PlusJacobian(const Scalar* x, Scalar* jacobian /* ambient x tangent */) const;
  bool RightMultiplyByPlusJacobian(const Scalar* x, int num_rows, const Scalar* ambient_matrix, Scalar* tangent_matrix) const;
  bool Minus(const Scalar* y, const Scalar* x, Scalar* y_minus_x) const;
  bool MinusJacobian(const Scalar* x, Scalar* jacobian /* tangent x ambient */) const;

 private:
  explicit Manifold(std::vector<Part> parts) : parts_{std::move(parts)} {}
  static int tangent(const Part& p) { return p.kind == kEuclidean ? p.ambient : (p.kind == kConstant ? 0 : p.ambient - 1); }
  std::vector<Part> parts_;
};

}  // namespace hyper
