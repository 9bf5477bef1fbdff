// hyper::state -- ContinuousState / TemporalInterpolator (north_star names; the reference at this
// commit spells them AbstractState / BasisInterpolator, reference internal/hyper/optimizers/abstract.cpp:79-81).
// The state owns its control points; interpolation itself runs on the device through the C-ABI.
#pragma once
#include <algorithm>
#include <memory>
#include <stdexcept>
#include <vector>

#include "hyper/variables.hpp"

struct hb200_ctx;

namespace hyper {

struct TemporalInterpolatorLayout {
  struct { Index size; } outer;
  struct Padding { Index left, right; };
  Padding outerPadding() const { return {(outer.size - 1) / 2, outer.size - 1 - (outer.size - 1) / 2}; }
};

class TemporalInterpolator {
 public:
  virtual ~TemporalInterpolator() = default;
  virtual TemporalInterpolatorLayout layout() const = 0;
};

// BasisInterpolator(degree, uniform) (reference tests/internal/tests/optimizers/evaluators/pixel.cpp:50).
class BasisInterpolator final : public TemporalInterpolator {
 public:
  explicit BasisInterpolator(int degree = 3, bool uniform = true) : degree_{degree} {
    if (!uniform) throw std::invalid_argument("only uniform B-splines are supported");
    if (degree != 3 && degree != 5) throw std::invalid_argument("degree must be 3 or 5");
  }
  TemporalInterpolatorLayout layout() const override { return {{degree_ + 1}}; }
  int degree() const { return degree_; }
 private:
  int degree_;
};

struct StateQuery { Stamp stamp; Index derivative = 0; bool jacobian = false; };
struct StateResult { SE3 value; Tangent6 velocity, acceleration; };

template <typename T> struct Range { T lower, upper; bool contains(const T& t) const { return lower <= t && t < upper; } };

class ContinuousState {
 public:
  using Element = Stamped<SE3>;
  explicit ContinuousState(std::unique_ptr<TemporalInterpolator> interpolator = std::make_unique<BasisInterpolator>())
      : interpolator_{std::move(interpolator)} {}
  std::vector<Element>& elements() { return elements_; }            // kept ordered by stamp
  const std::vector<Element>& elements() const { return elements_; }
  const TemporalInterpolator* interpolator() const { return interpolator_.get(); }
  std::unique_ptr<TemporalInterpolator>& interpolator() { return interpolator_; }
  void sort() { std::sort(elements_.begin(), elements_.end(), [](const Element& a, const Element& b) { return a.stamp() < b.stamp(); }); }
  // Lower-inclusive range of stamps the spline can be evaluated at.
  Range<Stamp> range() const {
    const auto pad = interpolator_->layout().outerPadding();
    return {elements_[pad.left].stamp(), elements_[elements_.size() - 1 - pad.right].stamp()};
  }
  // Index of the first control point a stamp touches, -1 outside range().
  Index baseIndex(Stamp t) const {
    if (!range().contains(t)) return -1;
    auto it = std::upper_bound(elements_.begin(), elements_.end(), t, [](Stamp s, const Element& e) { return s < e.stamp(); });
    return static_cast<Index>(it - elements_.begin()) - 1 - interpolator_->layout().outerPadding().left;
  }
  // Parameter blocks of the control points a stamp touches (reference exteroceptive.cpp:35).
  Pointers<Scalar> parameters(Stamp t) {
    Pointers<Scalar> p;
    const Index b = baseIndex(t);
    if (b < 0) return p;
    for (Index i = 0; i < interpolator_->layout().outer.size; ++i) p.push_back(elements_[b + i].data());
    return p;
  }
  // Batched evaluation on the device (hb200_interpolate); derivative in {0, 2}.
  std::vector<StateResult> evaluate(hb200_ctx* ctx, const std::vector<Stamp>& stamps, Index derivative = 0) const;
 private:
  std::unique_ptr<TemporalInterpolator> interpolator_;
  std::vector<Element> elements_;
};

using AbstractState = ContinuousState;   // reference spelling at commit 5944de9

}  // namespace hyper
