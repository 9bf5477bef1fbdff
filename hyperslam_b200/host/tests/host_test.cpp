// GPU test of the C++ plugin surface: reads a flat window dump (written by tests/test_gpu_host.py),
// builds the hyper:: object graph the way the reference's optimizer holds it, and
//  (1) runs the reference's Probe protocol through ExteroceptiveCost::Evaluate + Manifold hooks
//      (analytic J_ambient * PlusJacobian vs central differences through Manifold::Plus, re-evaluated
//      on the device) -- reference tests/include/tests/optimizers/evaluators/evaluator.hpp:38-65;
//  (2) dumps residuals/Jacobians of a few costs and the state after Optimizer::optimize(5) for the
//      Python side to compare with the oracle.
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <memory>

#include "hyper/optimizer.hpp"

using namespace hyper;

static std::vector<double> readv(std::istream& in, size_t n) { std::vector<double> v(n); for (auto& x : v) in >> x; return v; }

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: host_test <window.txt> <out.txt>\n"); return 2; }
  std::ifstream in(argv[1]);
  if (!in) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
  int order, K, Kb, C, L, Nv, Ni, nconst, Nb, Nm, P;
  in >> order >> K >> Kb >> C >> L >> Nv >> Ni >> nconst >> Nb >> Nm >> P;
  ContinuousState state(std::make_unique<BasisInterpolator>(order - 1, true));
  for (int j = 0; j < K; ++j) { ContinuousState::Element e; for (int i = 0; i < 8; ++i) in >> e.v[i]; state.elements().push_back(e); }
  IMU imu;
  for (int j = 0; j < Kb; ++j) { IMU::Bias b; for (int i = 0; i < 4; ++i) in >> b.v[i]; imu.gyroscopeBias().push_back(b); }
  for (int j = 0; j < Kb; ++j) { IMU::Bias b; for (int i = 0; i < 4; ++i) in >> b.v[i]; imu.accelerometerBias().push_back(b); }
  Gravity gravity; for (int i = 0; i < 3; ++i) in >> gravity[i];
  std::vector<std::unique_ptr<Camera>> cams;
  for (int c = 0; c < C; ++c) {
    cams.emplace_back(new Camera());
    for (int i = 0; i < 7; ++i) in >> cams[c]->transformation().v[i];
    for (int i = 0; i < 4; ++i) in >> cams[c]->intrinsics()[i];
    for (int i = 0; i < 4; ++i) in >> cams[c]->distortion()[i];
  }
  { auto v = readv(in, 37); auto p = imu.variables(); const int s[5] = {7, 6, 6, 9, 9}; int o = 0; for (int b = 0; b < 5; ++b) { std::copy(v.begin() + o, v.begin() + o + s[b], p[b]); o += s[b]; } }
  std::vector<Landmark> lms(L);
  for (int l = 0; l < L; ++l) for (int i = 0; i < 3; ++i) in >> lms[l].variable[i];
  std::vector<VisualPixelObservation> vobs(Nv);
  for (int f = 0; f < Nv; ++f) { int c, l; in >> vobs[f].stamp >> c >> l >> vobs[f].measurement[0] >> vobs[f].measurement[1]; if (!in || c < 0 || c >= C || l < 0 || l >= L) { std::fprintf(stderr, "malformed pixel factor %d\n", f); return 2; } vobs[f].camera = cams[c].get(); vobs[f].landmark = &lms[l]; }
  std::vector<InertialObservation> iobs(Ni);
  for (int f = 0; f < Ni; ++f) { in >> iobs[f].stamp; for (int i = 0; i < 6; ++i) in >> iobs[f].measurement[i]; iobs[f].imu = &imu; iobs[f].gravity = &gravity; }
  std::vector<VisualBearingObservation> bobs(Nb);
  for (int f = 0; f < Nb; ++f) { int c, l; in >> bobs[f].stamp >> c >> l; for (int i = 0; i < 3; ++i) in >> bobs[f].measurement[i]; bobs[f].camera = cams[c].get(); bobs[f].landmark = &lms[l]; }
  std::vector<Sensor> pose_sensors(P);
  for (int p = 0; p < P; ++p) for (int i = 0; i < 7; ++i) in >> pose_sensors[p].transformation().v[i];
  std::vector<ManifoldObservation> mobs(Nm);
  for (int f = 0; f < Nm; ++f) { int sidx; in >> mobs[f].stamp >> sidx; for (int i = 0; i < 7; ++i) in >> mobs[f].measurement.v[i]; mobs[f].sensor = &pose_sensors[sidx]; }
  if (!in) { std::fprintf(stderr, "window file is truncated\n"); return 2; }

  if (!in) { std::fprintf(stderr, "malformed window file\n"); return 2; }
  Optimizer optimizer(0);
  optimizer.setState(&state);
  std::vector<Camera*> cam_ptrs; for (auto& c : cams) cam_ptrs.push_back(c.get());
  optimizer.setCameras(cam_ptrs);
  optimizer.setIMU(&imu);
  optimizer.setGravity(&gravity);
  for (auto& l : lms) optimizer.addLandmark(l);
  std::vector<ExteroceptiveCost*> vcosts, icosts, bcosts, mcosts;
  for (auto& o : vobs) vcosts.push_back(optimizer.add(o));
  for (auto& o : iobs) icosts.push_back(optimizer.add(o));
  for (auto& o : bobs) bcosts.push_back(optimizer.add(o));
  for (auto& o : mobs) mcosts.push_back(optimizer.add(o));
  std::vector<bool> constant(K, false); for (int j = 0; j < nconst; ++j) constant[j] = true;
  optimizer.setStateConstant(constant);

  std::ofstream out(argv[2]);
  out.precision(17);
  const Manifold state_manifold = Manifold::StampedSE3(true, false, false);   // reference pixel.cpp(test):52
  const Manifold landmark_manifold = Manifold::Euclidean(3);
  const Manifold bias_manifold = Manifold::StampedEuclidean(3, true, false);   // reference imu.cpp:64-66
  const Manifold gravity_manifold = Manifold::Sphere(3);

  // ---- (1) Probe through the plugin surface, on the variable blocks of a few costs ----
  int probe_failures = 0, probe_blocks = 0;
  auto probe = [&](ExteroceptiveCost* cost, const std::vector<std::pair<int, const Manifold*>>& blocks) {
    auto params = cost->update();
    const int nr = cost->num_residuals();
    const auto& sizes = cost->parameter_block_sizes();
    std::vector<std::vector<double>> J(params.size());
    std::vector<double*> jp(params.size());
    for (size_t b = 0; b < params.size(); ++b) { J[b].assign(static_cast<size_t>(nr) * sizes[b], 0.0); jp[b] = J[b].data(); }
    std::vector<double> r0(nr), rp(nr), rm(nr);
    optimizer.prepareForEvaluation(true);
    cost->Evaluate(params.data(), r0.data(), jp.data());
    for (auto [b, m] : blocks) {
      const int na = m->AmbientSize(), nt = m->TangentSize();
      std::vector<double> Ja(static_cast<size_t>(nr) * nt), Jn(static_cast<size_t>(nr) * nt), save(params[b], params[b] + na), delta(nt), xp(na);
      m->RightMultiplyByPlusJacobian(params[b], nr, J[b].data(), Ja.data());
      const double h = 1e-6;
      for (int c = 0; c < nt; ++c) {
        for (int s = 0; s < 2; ++s) {
          std::fill(delta.begin(), delta.end(), 0.0); delta[c] = s ? -h : h;
          m->Plus(save.data(), delta.data(), xp.data());
          std::copy(xp.begin(), xp.end(), params[b]);
          optimizer.prepareForEvaluation(false);                      // new evaluation point -> one device batch
          cost->Evaluate(params.data(), s ? rm.data() : rp.data(), nullptr);
        }
        std::copy(save.begin(), save.end(), params[b]);
        for (int r = 0; r < nr; ++r) Jn[r * nt + c] = (rp[r] - rm[r]) / (2 * h);
      }
      double na2 = 0, nn2 = 0, worst_rel = 0, worst_abs = 0;
      for (size_t i = 0; i < Ja.size(); ++i) { na2 += Ja[i] * Ja[i]; nn2 += Jn[i] * Jn[i]; }
      for (size_t i = 0; i < Ja.size(); ++i) {
        const double ae = std::fabs(Ja[i] - Jn[i]);
        double re = ae / std::max(std::fabs(Ja[i]), std::fabs(Jn[i]));
        if (Ja[i] == 0.0 || Jn[i] == 0.0) re = ae;
        worst_rel = std::max(worst_rel, re);
        worst_abs = std::max(worst_abs, std::fabs(Ja[i] / std::sqrt(na2) - Jn[i] / std::sqrt(nn2)));
      }
      ++probe_blocks;
      if (worst_rel > 1e-5 && worst_abs > 1e-5) { ++probe_failures; std::fprintf(stderr, "probe: cost kind %d index %d block %d rel %.3e abs %.3e\n", cost->kind(), cost->index(), b, worst_rel, worst_abs); }
    }
  };
  const int k = order;
  if (Nv) for (int f : {0, Nv / 2}) { std::vector<std::pair<int, const Manifold*>> blk; for (int m = 1; m + 1 < k; ++m) blk.push_back({m, &state_manifold}); blk.push_back({k + 3, &landmark_manifold}); probe(vcosts[f], blk); }
  if (Ni) for (int f : {Ni / 3}) { std::vector<std::pair<int, const Manifold*>> blk; for (int m = 1; m + 1 < k; ++m) blk.push_back({m, &state_manifold}); blk.push_back({k + 5 + 1, &bias_manifold}); blk.push_back({k + 5 + 4 + 2, &bias_manifold}); blk.push_back({k + 13, &gravity_manifold}); probe(icosts[f], blk); }
  if (Nb) for (int f : {0, Nb - 1}) { std::vector<std::pair<int, const Manifold*>> blk; for (int m = 1; m + 1 < k; ++m) blk.push_back({m, &state_manifold}); blk.push_back({k + 3, &landmark_manifold}); probe(bcosts[f], blk); }
  if (Nm) for (int f : {0, Nm - 1}) { std::vector<std::pair<int, const Manifold*>> blk; for (int m = 1; m + 1 < k; ++m) blk.push_back({m, &state_manifold}); probe(mcosts[f], blk); }
  out << "probe " << probe_blocks << " " << probe_failures << "\n";

  // ---- (2) copy-out of a few costs + state interpolation + optimize ----
  optimizer.prepareForEvaluation(true);
  auto dump_cost = [&](ExteroceptiveCost* cost) {
    auto params = cost->update();
    const int nr = cost->num_residuals();
    const auto& sizes = cost->parameter_block_sizes();
    std::vector<std::vector<double>> J(params.size());
    std::vector<double*> jp(params.size());
    for (size_t b = 0; b < params.size(); ++b) { J[b].assign(static_cast<size_t>(nr) * sizes[b], 0.0); jp[b] = J[b].data(); }
    std::vector<double> r(nr);
    cost->Evaluate(params.data(), r.data(), jp.data());
    out << "cost " << cost->kind() << " " << cost->index() << " " << nr << " " << cost->layout().num_parameters << " " << params.size();
    for (int s : sizes) out << " " << s;
    out << "\n";
    for (double x : r) out << x << " ";
    out << "\n";
    for (auto& Jb : J) { for (double x : Jb) out << x << " "; out << "\n"; }
  };
  if (Nv) { dump_cost(vcosts[0]); dump_cost(vcosts[Nv - 1]); }
  if (Ni) { dump_cost(icosts[0]); dump_cost(icosts[Ni - 1]); }
  if (Nb) { dump_cost(bcosts[0]); dump_cost(bcosts[Nb - 1]); }
  if (Nm) { dump_cost(mcosts[0]); dump_cost(mcosts[Nm - 1]); }
  const auto range = state.range();
  std::vector<Stamp> ts; for (int i = 0; i < 5; ++i) ts.push_back(range.lower + (range.upper - range.lower) * (i + 0.5) / 5.0);
  auto sr = state.evaluate(optimizer.context(), ts, 2);
  out << "interp " << ts.size() << "\n";
  for (size_t i = 0; i < ts.size(); ++i) { out << ts[i]; for (double x : sr[i].value.v) out << " " << x; for (double x : sr[i].velocity.v) out << " " << x; for (double x : sr[i].acceleration.v) out << " " << x; out << "\n"; }
  auto summary = optimizer.optimize(5);
  out << "optimize " << summary.size() << "\n";
  for (auto& s : summary) out << s.cost << " " << s.cost_new << " " << s.accepted << " " << s.spd << "\n";
  out << "knots\n";
  for (auto& e : state.elements()) { for (double x : e.v) out << x << " "; out << "\n"; }
  out << "landmarks\n";
  for (auto& l : lms) { for (double x : l.variable.v) out << x << " "; out << "\n"; }
  out << "gravity " << gravity[0] << " " << gravity[1] << " " << gravity[2] << "\n";
  std::printf("host_test: probe blocks %d failures %d; cost %.6f -> %.6f\n", probe_blocks, probe_failures, summary.front().cost, summary.back().cost_new);
  return probe_failures ? 1 : 0;
}
