// Shared host/device declarations for libhyperb200 (sm_100a).
#pragma once
#include <cstdint>

namespace hb {

constexpr int kMaxOrder = 6;
// Knot-table row (one per control point, rebuilt once per evaluation by prep_kernel):
//   [0..8]  R_j        rotation matrix of control point j (row-major)
//   [9..11] p_j
//   [12..14] d_j = Log(R_{j-1}^T R_j)                (row 0: zeros)
//   [15..23] G_j = Jr^{-1}(d_j) R_j^T                (row 0: zeros)
//   [24]    stamp_j ; [25..27] pad  -> 28 doubles = 224 B (16 B multiple for cp.async.bulk)
constexpr int kTabStride = 28;
constexpr int kTileRows = 16;       // knot-table rows staged per CTA (3.6 KB); wider spans read the table from global/L2
constexpr int kEvalThreads = 64;    // one factor per thread

constexpr int kCamStride = 20;      // R_sb(9) t_bs(3) intrinsics(4) distortion(4)
// IMU derived table: R_sb(9) t_bs(3) IgRsb(9) IaRsb(9) Sg(9,row-major) lever arms c_r (3x3 rows) = 48
constexpr int kImuStride = 48;

struct Basis {
  int k;
  double Mc[kMaxOrder * kMaxOrder];  // cumulative blending matrix, row j = coefficients of lambda_j(u)
};

struct SolverState {  // lives on the device; updated by accept_kernel
  double radius;
  double decrease_factor;
  double cost;        // cost at the current linearisation point
  double cost_new;
  double model_change;
  double rho;
  int accepted;
  int spd;
  int iteration;
  int pad;
};

}  // namespace hb
