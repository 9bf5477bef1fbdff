// hyper::variables -- minimal host-side variable types with the storage the reference uses
// (HyperVariables is not available offline; names and memory layouts follow its call sites):
//   SU2 = [qx qy qz qw], SE3 = [q | p], Stamped<T> = [T | stamp]
//   (reference resources/.../settings.yaml:34-36, include/hyper/optimizers/ceres/manifolds/variables/stamped.hpp:35-36).
// Parameter memory is owned by the variables and aliased by pointer, exactly as the reference does
// with Ceres (reference internal/hyper/optimizers/ceres/optimizer.cpp:151,305,354-356).
#pragma once
#include <array>
#include <cstddef>
#include <vector>

namespace hyper {

using Scalar = double;
using Index = int;
using Stamp = double;
template <typename T>
using Pointers = std::vector<T*>;

template <int N>
struct Cartesian {
  std::array<Scalar, N> v{};
  Scalar* data() { return v.data(); }
  const Scalar* data() const { return v.data(); }
  static constexpr int kNumParameters = N;
  Scalar& operator[](int i) { return v[i]; }
  const Scalar& operator[](int i) const { return v[i]; }
};

using Position = Cartesian<3>;
using Pixel = Cartesian<2>;
using Gravity = Cartesian<3>;   // lives on the sphere |g| = const (reference environment/abstract.cpp:59-65)

struct SU2 : Cartesian<4> {
  SU2() { v = {0, 0, 0, 1}; }
};

struct SE3 : Cartesian<7> {
  SE3() { v = {0, 0, 0, 1, 0, 0, 0}; }
  Scalar* rotation() { return v.data(); }
  Scalar* translation() { return v.data() + 4; }
  const Scalar* rotation() const { return v.data(); }
  const Scalar* translation() const { return v.data() + 4; }
};

// [variable | stamp]
template <typename T>
struct Stamped {
  std::array<Scalar, T::kNumParameters + 1> v{};
  static constexpr int kNumParameters = T::kNumParameters + 1;
  Stamped() { T t; for (int i = 0; i < T::kNumParameters; ++i) v[i] = t.v[i]; }
  Scalar* data() { return v.data(); }
  const Scalar* data() const { return v.data(); }
  Stamp& stamp() { return v[T::kNumParameters]; }
  const Stamp& stamp() const { return v[T::kNumParameters]; }
};

struct Tangent6 : Cartesian<6> {   // Tangent<SE3>: [angular | linear] (reference inertial.cpp:135-150)
  Scalar* angular() { return v.data(); }
  Scalar* linear() { return v.data() + 3; }
};

}  // namespace hyper
