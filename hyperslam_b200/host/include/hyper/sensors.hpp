// hyper::sensors -- Camera / IMU parameter holders with the block order of Sensor::parameters()
// (reference pixel.cpp:43-45 Traits<Camera>::{kTransformation,kIntrinsics,kDistortion}Offset;
//  inertial.cpp:35-39 Traits<IMU>::k...Offset).
#pragma once
#include <vector>

#include "hyper/variables.hpp"

namespace hyper {

class Sensor {
 public:
  virtual ~Sensor() = default;
  SE3& transformation() { return T_bs_; }
  const SE3& transformation() const { return T_bs_; }
  virtual Pointers<Scalar> variables() { return {T_bs_.data()}; }
 protected:
  SE3 T_bs_;
};

class Camera final : public Sensor {
 public:
  static constexpr int kTransformationOffset = 0, kIntrinsicsOffset = 1, kDistortionOffset = 2, kNumParameters = 3;
  Cartesian<4>& intrinsics() { return intrinsics_; }    // [cx cy fx fy] (reference settings.yaml:38-40)
  Cartesian<4>& distortion() { return distortion_; }    // radtan [k1 k2 p1 p2] (settings.yaml:42-45)
  const Cartesian<4>& intrinsics() const { return intrinsics_; }
  const Cartesian<4>& distortion() const { return distortion_; }
  Pointers<Scalar> variables() override { return {T_bs_.data(), intrinsics_.data(), distortion_.data()}; }
 private:
  Cartesian<4> intrinsics_, distortion_;
};

class IMU final : public Sensor {
 public:
  static constexpr int kTransformationOffset = 0, kGyroscopeIntrinsicsOffset = 1, kAccelerometerIntrinsicsOffset = 2,
                       kGyroscopeSensitivityOffset = 3, kAccelerometerAxesOffsetsOffset = 4, kNumParameters = 5;
  using Bias = Stamped<Cartesian<3>>;   // [bx by bz | stamp]
  IMU() { gyro_intrinsics_.v = {1, 1, 1, 0, 0, 0}; accel_intrinsics_.v = {1, 1, 1, 0, 0, 0}; }
  Cartesian<6>& gyroscopeIntrinsics() { return gyro_intrinsics_; }
  Cartesian<6>& accelerometerIntrinsics() { return accel_intrinsics_; }
  Cartesian<9>& gyroscopeSensitivity() { return S_g_; }
  Cartesian<9>& accelerometerAxesOffsets() { return X_a_; }
  std::vector<Bias>& gyroscopeBias() { return gyro_bias_; }         // bias spline control points (order 4)
  std::vector<Bias>& accelerometerBias() { return accel_bias_; }
  Pointers<Scalar> variables() override {
    return {T_bs_.data(), gyro_intrinsics_.data(), accel_intrinsics_.data(), S_g_.data(), X_a_.data()};
  }
 private:
  Cartesian<6> gyro_intrinsics_, accel_intrinsics_;
  Cartesian<9> S_g_, X_a_;
  std::vector<Bias> gyro_bias_, accel_bias_;
};

}  // namespace hyper
