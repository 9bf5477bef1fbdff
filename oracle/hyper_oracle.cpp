// TEST INFRASTRUCTURE ONLY -- C entry points of the CPU oracle (loaded with ctypes by tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs).
// PARITY UNPINNED (see ho_math.h).  Never linked into libhyperb200.so.
#include <algorithm>
#include <cstdio>
#include <vector>

#include "ho_evaluators.h"
#include "ho_manifolds.h"
#include "ho_ingest.h"
#include "ho_window.h"

using namespace ho;

namespace {

// Block manifold ids used by the probe / host tests.
enum BlockManifold { kMState = 0, kMSE3 = 1, kMEuclidean = 2, kMConstant = 3, kMBias = 4, kMSphere = 5 };

Manifold make_block_manifold(int id, int size) {
  switch (id) {
    case kMState: return make_stamped_se3(true, false, false);  // reference pixel.cpp(test):52 {true,false,false}
    case kMSE3: return make_se3(false, false);
    case kMEuclidean: return make_cartesian(size, false);
    case kMConstant: return make_cartesian(size, true);
    case kMBias: return make_stamped_cartesian(size - 1, true, false);  // reference imu.cpp:64-66
    case kMSphere: return make_sphere(size, false);
  }
  return make_cartesian(size, false);
}

Factor make_factor(int kind, double stamp, const double* meas, int k, int k_bg, int k_ba) {
  Factor f;
  f.kind = kind; f.stamp = stamp; f.k = k; f.k_bg = k_bg; f.k_ba = k_ba;
  const int nm = measurement_size(f);
  for (int i = 0; i < nm; ++i) f.measurement[i] = meas[i];
  return f;
}

}  // namespace

extern "C" {

int ho_basis(int k, double* Mc) {
  if (k < 2 || k > kMaxOrder) return -1;
  Basis b; basis_init(&b, k);
  for (int j = 0; j < k; ++j) for (int n = 0; n < k; ++n) Mc[j * k + n] = b.Mc[j][n];
  return 0;
}

int ho_basis_eval(int k, double u, double inv_dt, double* lam /*3 x k*/) {
  Basis b; basis_init(&b, k);
  double l[3][kMaxOrder + 1];
  basis_eval(b, u, inv_dt, l);
  for (int d = 0; d < 3; ++d) for (int j = 0; j < k; ++j) lam[d * k + j] = l[d][j];
  return 0;
}

// cps: k x 8 contiguous.  J: 3 x 6 x 8k (only the first derivative+1 slabs are written).
int ho_state_evaluate(int k, const double* cps, double stamp, int derivative, int jac, double* value, double* velocity,
                      double* acceleration, double* J) {
  Basis b; basis_init(&b, k);
  const double* ptrs[kMaxOrder];
  for (int i = 0; i < k; ++i) ptrs[i] = cps + 8 * i;
  StateResult S;
  state_evaluate(b, ptrs, stamp, derivative, jac != 0, &S);
  for (int i = 0; i < 7; ++i) value[i] = S.value[i];
  if (derivative >= 1) for (int i = 0; i < 6; ++i) { velocity[i] = S.velocity[i]; acceleration[i] = S.acceleration[i]; }
  if (jac) {
    const int nd = derivative >= 2 ? 3 : 1;
    for (int d = 0; d < nd; ++d) std::memcpy(J + (size_t)d * 6 * 8 * k, S.J[d], sizeof(double) * 6 * 8 * k);
  }
  return 0;
}

// out: [num_blocks, num_parameters, state_idx, sensor_static_idx, sensor_dynamic_idx, observation_idx],
// offsets[num_blocks], sizes[num_blocks]
int ho_layout(int kind, int k, int k_bg, int k_ba, int* out, int* offsets, int* sizes) {
  Factor f; f.kind = kind; f.k = k; f.k_bg = k_bg; f.k_ba = k_ba;
  Layout L; layout_update(f, &L);
  out[0] = L.num_blocks; out[1] = L.num_parameters; out[2] = L.state_idx; out[3] = L.sensor_static_idx;
  out[4] = L.sensor_dynamic_idx; out[5] = L.observation_idx;
  for (int i = 0; i < L.num_blocks; ++i) { offsets[i] = L.offsets[i]; sizes[i] = L.sizes[i]; }
  return 0;
}

// Per-factor reference-shaped Evaluate.  params: blocks concatenated in layout order.
// jac_mask: per block 0/1 (null => all blocks); jac (may be null => cost only): per-block row-major
// (nr x size) slabs concatenated at nr * offsets[b].
int ho_cost_evaluate(int kind, double stamp, const double* meas, int k, int k_bg, int k_ba, const double* params,
                     const int* jac_mask, double* residuals, double* jac, int quirks) {
  Factor f = make_factor(kind, stamp, meas, k, k_bg, k_ba);
  Layout L; layout_update(f, &L);
  Basis b, bb; basis_init(&b, k); basis_init(&bb, k_bg);
  const double* ptrs[kMaxBlocks];
  double* jptrs[kMaxBlocks];
  const int nr = num_residuals(f);
  for (int i = 0; i < L.num_blocks; ++i) {
    ptrs[i] = params + L.offsets[i];
    jptrs[i] = (jac && (!jac_mask || jac_mask[i])) ? jac + (size_t)nr * L.offsets[i] : nullptr;
  }
  return cost_evaluate(f, L, b, bb, ptrs, residuals, jac ? jptrs : nullptr, quirks) ? 0 : 1;
}

// Restatement of EvaluatorTests<CERES>::Probe (reference
// tests/include/tests/optimizers/evaluators/evaluator.hpp:38-65): analytic local Jacobian
// (ambient Jacobian x PlusJacobian) vs central differences through the manifold Plus.
// manifold_ids: per block BlockManifold.  results: [relative_ok, absolute_ok, max_rel_err, max_abs_norm_err]
// per_block_err (optional): num_blocks x 2.
int ho_probe(int kind, double stamp, const double* meas, int k, int k_bg, int k_ba, const double* params,
             const int* manifold_ids, double tolerance, int quirks, double* results, double* per_block_err, int richardson) {
  Factor f = make_factor(kind, stamp, meas, k, k_bg, k_ba);
  Layout L; layout_update(f, &L);
  Basis b, bb; basis_init(&b, k); basis_init(&bb, k_bg);
  const int nr = num_residuals(f);
  std::vector<double> x(params, params + L.num_parameters);
  const double* ptrs[kMaxBlocks];
  double* jptrs[kMaxBlocks];
  std::vector<double> jac((size_t)nr * L.num_parameters);
  for (int i = 0; i < L.num_blocks; ++i) { ptrs[i] = x.data() + L.offsets[i]; jptrs[i] = jac.data() + (size_t)nr * L.offsets[i]; }
  double r0[6];
  cost_evaluate(f, L, b, bb, ptrs, r0, jptrs, quirks);
  bool rel_ok = true, abs_ok = true;
  double max_rel = 0, max_abs = 0;
  const double h = 1e-6;  // relative_step_size / ridders initial step, evaluator.hpp:28-29
  for (int blk = 0; blk < L.num_blocks; ++blk) {
    Manifold m = make_block_manifold(manifold_ids[blk], L.sizes[blk]);
    const int na = m.ambient_size(), nt = m.tangent_size();
    double blk_rel = 0, blk_abs = 0;
    if (nt > 0) {
      std::vector<double> PJ(na * nt), Ja(nr * nt), Jn(nr * nt);
      manifold_plus_jacobian(m, x.data() + L.offsets[blk], PJ.data());
      mat_mul(jptrs[blk], PJ.data(), Ja.data(), nr, na, nt);
      std::vector<double> save(x.begin() + L.offsets[blk], x.begin() + L.offsets[blk] + na), delta(nt), xp(na);
      // richardson = 0: plain central difference with the reference's step (evaluator.hpp:26-32);
      // richardson = 1: one Richardson extrapolation from steps 2e-4 / 1e-4 (O(h^4), ~100x less
      // round-off) -- used where a block's Jacobian is tiny (edge control points of a quintic spline).
      for (int c = 0; c < nt; ++c) {
        auto central = [&](double step, double* out) {
          double rp[6], rm[6];
          std::fill(delta.begin(), delta.end(), 0.0);
          delta[c] = step;
          manifold_plus(m, save.data(), delta.data(), xp.data());
          std::copy(xp.begin(), xp.end(), x.begin() + L.offsets[blk]);
          cost_evaluate(f, L, b, bb, ptrs, rp, nullptr, quirks);
          delta[c] = -step;
          manifold_plus(m, save.data(), delta.data(), xp.data());
          std::copy(xp.begin(), xp.end(), x.begin() + L.offsets[blk]);
          cost_evaluate(f, L, b, bb, ptrs, rm, nullptr, quirks);
          for (int r = 0; r < nr; ++r) out[r] = (rp[r] - rm[r]) / (2 * step);
        };
        double d1[6], d2[6];
        if (!richardson) {
          central(h, d1);
          for (int r = 0; r < nr; ++r) Jn[r * nt + c] = d1[r];
        } else {
          central(2e-4, d1);
          central(1e-4, d2);
          for (int r = 0; r < nr; ++r) Jn[r * nt + c] = (4.0 * d2[r] - d1[r]) / 3.0;
        }
      }
      std::copy(save.begin(), save.end(), x.begin() + L.offsets[blk]);
      double na2 = 0, nn2 = 0;
      for (int i = 0; i < nr * nt; ++i) { na2 += Ja[i] * Ja[i]; nn2 += Jn[i] * Jn[i]; }
      const double ina = na2 > 0 ? 1.0 / std::sqrt(na2) : 0.0, inn = nn2 > 0 ? 1.0 / std::sqrt(nn2) : 0.0;
      for (int i = 0; i < nr * nt; ++i) {
        const double a = Ja[i], n = Jn[i];
        const double ae = std::fabs(a - n);
        double re = ae / std::max(std::fabs(a), std::fabs(n));
        if (a == 0.0 || n == 0.0) re = ae;
        blk_rel = std::max(blk_rel, re);
        blk_abs = std::max(blk_abs, std::fabs(a * ina - n * inn));
      }
    }
    if (blk_rel > tolerance) rel_ok = false;
    if (blk_abs > tolerance) abs_ok = false;
    max_rel = std::max(max_rel, blk_rel);
    max_abs = std::max(max_abs, blk_abs);
    if (per_block_err) { per_block_err[2 * blk] = blk_rel; per_block_err[2 * blk + 1] = blk_abs; }
  }
  results[0] = rel_ok; results[1] = abs_ok; results[2] = max_rel; results[3] = max_abs;
  return (rel_ok || abs_ok) ? 0 : 1;
}

// Manifold hooks (reference wrapper.hpp:24-50) on a block-manifold id.
int ho_manifold_sizes(int id, int size, int* ambient, int* tangent) {
  Manifold m = make_block_manifold(id, size);
  *ambient = m.ambient_size(); *tangent = m.tangent_size();
  return 0;
}
int ho_manifold_plus(int id, int size, const double* x, const double* delta, double* out) {
  manifold_plus(make_block_manifold(id, size), x, delta, out);
  return 0;
}
int ho_manifold_plus_jacobian(int id, int size, const double* x, double* J) {
  manifold_plus_jacobian(make_block_manifold(id, size), x, J);
  return 0;
}
int ho_manifold_minus(int id, int size, const double* y, const double* x, double* out) {
  manifold_minus(make_block_manifold(id, size), y, x, out);
  return 0;
}

}  // extern "C"

extern "C" {
// Stereo-frame ingest (ho_ingest.h): cps = the k control points around `stamp` ([k][8]).
int ho_ingest_stereo_frame(int k, const double* cps, double stamp, const double* cam0, const double* cam1, int n, const double* px0,
                           const double* px1, double* B0, double* B1, double* landmarks) {
  Basis b; basis_init(&b, k);
  const double* ptrs[kMaxOrder];
  for (int i = 0; i < k; ++i) ptrs[i] = cps + 8 * i;
  ingest_stereo_frame(b, ptrs, stamp, cam0, cam1, n, px0, px1, B0, B1, landmarks);
  return 0;
}
}  // extern "C"

#include "ho_window_api.inc"
