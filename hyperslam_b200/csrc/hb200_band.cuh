// Block-banded + arrowhead Cholesky solve of the reduced system (sm_100a), one CTA.
//
// After the landmark Schur complement the reduced system of a spline window is
//   S = [ P  A^T ]   P: 6x6 control-point blocks, block half-bandwidth beta (= longest landmark track
//       [ A  C   ]      measured in control points, >= k-1), A: m x 6K arrow (bias knots + gravity),
//   C: m x m.  The reference hands the same structure to CHOLMOD through SPARSE_NORMAL_CHOLESKY
//   (reference internal/hyper/optimizers/ceres/optimizer.cpp:46-48); here the factorisation,
//   forward and backward substitution run in a single CTA with the whole band resident in shared
//   memory (global-memory workspace when it does not fit).
//
// Workspace (doubles): W[K][h][6] band block-columns (h = 6 + 6 beta rows each: diagonal block first),
//   AR[m+1][np] arrow rows (+ the rhs as last row), CC[m+1][m] corner (+ rhs row), LI[K][48] reciprocal
//   diagonals (6) and inverses (36, at +8) of the diagonal Cholesky blocks, X[n].
#pragma once
#include "hb200_solve.cuh"

namespace hb {

constexpr int kBandThreads = 512;

__host__ __device__ inline size_t band_workspace_doubles(int K, int beta, int m) {
  const size_t h = 6 + 6 * static_cast<size_t>(beta);
  const size_t np = 6 * static_cast<size_t>(K);
  return static_cast<size_t>(K) * h * 6 + (m + 1) * np + static_cast<size_t>(m + 1) * m + static_cast<size_t>(K) * 48 + np + m;
}

// Right-looking Cholesky of a 6x6 SPD block (lower, row-major with stride ld), division-free:
// inv[j] = 1 / L[j][j] comes straight out of rsqrt.  If X != nullptr the block first receives its
// pending trailing update A -= X X^T (X = the 6 panel rows of this block, row-major 6x6), so the
// look-ahead thread does not wait for a separate update pass.
HB_DI bool chol6(double* A, int ld, double* inv /*6*/, const double* X = nullptr) {
  double L[21];  // packed lower: L[i(i+1)/2 + j]
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) L[i * (i + 1) / 2 + j] = A[i * ld + j];
  if (X != nullptr) {
    double x[36];
#pragma unroll
    for (int e = 0; e < 18; ++e) { const double2 t = reinterpret_cast<const double2*>(X)[e]; x[2 * e] = t.x; x[2 * e + 1] = t.y; }
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < 6; ++q) acc += x[6 * i + q] * x[6 * j + q];
        L[i * (i + 1) / 2 + j] -= acc;
      }
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const double d = L[j * (j + 1) / 2 + j];
    if (!(d > 0.0)) ok = false;
    const double iv = rsqrt(d);
    inv[j] = iv;
    L[j * (j + 1) / 2 + j] = d * iv;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) L[i * (i + 1) / 2 + j] *= iv;
#pragma unroll
    for (int i = j + 1; i < 6; ++i)
#pragma unroll
      for (int q = j + 1; q <= i; ++q) L[i * (i + 1) / 2 + q] -= L[i * (i + 1) / 2 + j] * L[q * (q + 1) / 2 + j];
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) A[i * ld + j] = L[i * (i + 1) / 2 + j];
  return ok;
}

template <bool SMEM>
__global__ void __launch_bounds__(kBandThreads) band_solve_kernel(const double* __restrict__ sys, int n, int K, int beta,
                                                                  double* __restrict__ ws_global, double* __restrict__ x_out,
                                                                  int* __restrict__ spd_flag, long long* __restrict__ dbg) {
  extern __shared__ double s_band[];
  double* ws = SMEM ? s_band : ws_global;
  const int np = 6 * K, m = n - np, h = 6 + 6 * beta;
  double* W = ws;
  double* AR = W + static_cast<size_t>(K) * h * 6;
  double* CC = AR + static_cast<size_t>(m + 1) * np;
  double* LI = CC + static_cast<size_t>(m + 1) * m;
  double* X = LI + static_cast<size_t>(K) * 48;
  const double* S = sys;
  const double* b = sys + static_cast<size_t>(n) * n;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  __shared__ int s_ok;
  __shared__ long long s_dbg[2], s_mark;   // look-ahead warp timing (debug)
  if (tid == 0) { s_ok = 1; s_dbg[0] = 0; s_dbg[1] = 0; }
  // ---- gather the band, the arrow and the corner from the dense system ----
  {
    const int h6 = h * 6;
    for (int e = tid; e < K * h6; e += kBandThreads) {
      const int c = e / h6, rem = e - c * h6;
      const int i = rem / 6, j = rem - 6 * i;
      const int row = 6 * c + i, col = 6 * c + j;
      W[e] = (row < np) ? S[static_cast<size_t>(row) * n + col] : 0.0;
    }
    for (int e = tid; e < (m + 1) * np; e += kBandThreads) {
      const int r = e / np, col = e - r * np;
      AR[e] = (r < m) ? S[static_cast<size_t>(np + r) * n + col] : b[col];
    }
    for (int e = tid; e < (m + 1) * m; e += kBandThreads) {
      const int r = e / m, q = e - r * m;
      CC[e] = (r < m) ? S[static_cast<size_t>(np + r) * n + np + q] : b[np + q];
    }
  }
  __syncthreads();
  long long t_mark = clock64(), t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define HB_TICK(i) do { if (dbg && tid == 0) { const long long now = clock64(); t_acc[i] += now - t_mark; t_mark = now; } } while (0)
  HB_TICK(0);
  // ---- factorisation + forward substitution, one control-point block column per step.  The 6x6
  // Cholesky of the NEXT diagonal block is done by warp 0 inside the trailing-update phase, as soon
  // as that block has received its update (look-ahead), so it is off the critical path. ----
  if (tid == 0 && K > 0) {
    if (!chol6(W, 6, LI)) s_ok = 0;
  }
  __syncthreads();
  HB_TICK(1);
  const int ng = (m + 1 + 5) / 6;
  for (int c = 0; c < K; ++c) {
    double* Wc = W + static_cast<size_t>(c) * h * 6;
    const double* Lic = LI + static_cast<size_t>(c) * 48;
    const int nb = min(h - 6, np - 6 * (c + 1));  // band rows below the diagonal block
    const int R = nb + m + 1;                      // + arrow rows + rhs row
    // panel: solve x L^T = a for every row (right-looking, reciprocal diagonal in Lic)
    for (int t = tid; t < R; t += kBandThreads) {
      double* a = (t < nb) ? (Wc + static_cast<size_t>(6 + t) * 6) : (AR + static_cast<size_t>(t - nb) * np + 6 * c);
      double v[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) v[q] = a[q];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        v[j] *= Lic[j];
#pragma unroll
        for (int q = j + 1; q < 6; ++q) v[q] -= v[j] * Wc[q * 6 + j];
      }
#pragma unroll
      for (int q = 0; q < 6; ++q) a[q] = v[q];
    }
    __syncthreads();
    HB_TICK(2);
    if (dbg && tid == 0) s_mark = t_mark;   // (racy by a few cycles with the readers; debug only)
    // trailing update in 6x6 tiles: groups = band blocks below the diagonal, then arrow rows in sixes
    // (the rhs row is the last arrow row).  Thread = (tile, row i of the tile): 6 dots of length 6.
    {
      const int nbk = nb / 6;
      const int G = nbk + ng;
      const int ntiles = G * (G + 1) / 2;
      const bool look = (c + 1 < K);   // then tile 0 is the next diagonal block and belongs to warp 0
      // warp 0 handles tile 0 only (when looking ahead) and keeps scheduler 0 to itself: the other
      // tiles go to the warps of schedulers 1..3 (warp % 4 != 0), so the critical path tile 0 -> chol6
      // does not compete for issue slots.
      int t;
      int stride;
      if (look) {
        // Scheduler 0 (warps 0, 4, 8, 12; the arbiter favours the highest warp id) is reserved for the
        // latency-critical look-ahead: warp 12 runs it, warp 8 hosts the block-inverse thread (lane 0)
        // and takes a few left-over items, warps 0 and 4 idle.  The 12 warps of schedulers 1..3 share
        // tiles 1 .. ntiles-1 (tile 0 = next diagonal block, done by the look-ahead warp).
        constexpr int kWorkers = 12 * 32 + 31;
        if ((warp & 3) != 0) { t = 6 + (warp - (warp >> 2) - 1) * 32 + lane; stride = kWorkers; }
        else if (warp == 8 && lane > 0) { t = 6 + 12 * 32 + lane - 1; stride = kWorkers; }
        else { t = ntiles * 6; stride = 1; }
      } else { t = tid; stride = kBandThreads; }
      for (; t < ntiles * 6; t += stride) {
        const int tile = t / 6, i = t - 6 * tile;
        int gu = static_cast<int>((sqrtf(8.0f * tile + 1.0f) - 1.0f) * 0.5f);
        while (gu * (gu + 1) / 2 > tile) --gu;
        while ((gu + 1) * (gu + 2) / 2 <= tile) ++gu;
        const int gv = tile - gu * (gu + 1) / 2;
        const bool ub = gu < nbk, vb = gv < nbk;
        const int ru = ub ? 0 : 6 * (gu - nbk) + i;   // arrow row index of u
        if (!ub && ru > m) continue;
        const double* xu = ub ? (Wc + static_cast<size_t>(6 + 6 * gu + i) * 6) : (AR + static_cast<size_t>(ru) * np + 6 * c);
        double a[6], sd[6];
        {
          const double2* x2 = reinterpret_cast<const double2*>(xu);   // rows are 48 B: 16-byte aligned
          const double2 a0 = x2[0], a1 = x2[1], a2 = x2[2];
          a[0] = a0.x; a[1] = a0.y; a[2] = a1.x; a[3] = a1.y; a[4] = a2.x; a[5] = a2.y;
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const int rv = vb ? 0 : 6 * (gv - nbk) + j;
          const double* xv = vb ? (Wc + static_cast<size_t>(6 + 6 * gv + j) * 6) : (AR + static_cast<size_t>(min(rv, m)) * np + 6 * c);
          const double2* v2 = reinterpret_cast<const double2*>(xv);
          const double2 b0 = v2[0], b1 = v2[1], b2 = v2[2];
          sd[j] = a[0] * b0.x + a[1] * b0.y + a[2] * b1.x + a[3] * b1.y + a[4] * b2.x + a[5] * b2.y;
        }
        if (ub) {  // band x band
          double* tgt = W + (static_cast<size_t>(c + 1 + gv) * h + (6 * (gu - gv) + i)) * 6;
#pragma unroll
          for (int j = 0; j < 6; ++j) tgt[j] -= sd[j];
        } else if (vb) {  // arrow x band
          double* tgt = AR + static_cast<size_t>(ru) * np + 6 * (c + 1 + gv);
#pragma unroll
          for (int j = 0; j < 6; ++j) tgt[j] -= sd[j];
        } else {  // arrow x arrow (columns are arrow dofs < m)
          double* tgt = CC + static_cast<size_t>(ru) * m + 6 * (gv - nbk);
#pragma unroll
          for (int j = 0; j < 6; ++j)
            if (6 * (gv - nbk) + j < m) tgt[j] -= sd[j];
        }
      }
      if (tid == 8 * 32) {   // lane 0 of warp 8
        // inverse of this step's diagonal factor, for the back substitution (off the critical path):
        // Li[i][j] = -(sum_{p=j}^{i-1} L[i][p] Li[p][j]) / L[i][i],  Li[j][j] = 1 / L[j][j]
        double* Li = LI + static_cast<size_t>(c) * 48 + 8;
#pragma unroll
        for (int jj = 0; jj < 6; ++jj) {
          Li[jj * 6 + jj] = Lic[jj];
#pragma unroll
          for (int ii = jj + 1; ii < 6; ++ii) {
            double acc = 0.0;
#pragma unroll
            for (int pp = jj; pp < ii; ++pp) acc -= Wc[ii * 6 + pp] * Li[pp * 6 + jj];
            Li[ii * 6 + jj] = acc * Lic[ii];
          }
        }
      }
      if (look && warp == 12) {
        // next diagonal block: lanes 0..5 apply row i of its update A -= X X^T from this step's panel
        // rows (tile 0), then lane 0 factors it
        double* An = W + static_cast<size_t>(c + 1) * h * 6;
        if (lane < 6) {
          const double2* xi = reinterpret_cast<const double2*>(Wc + 36 + 6 * lane);
          const double2 a0 = xi[0], a1 = xi[1], a2 = xi[2];
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            if (j <= lane) {
              const double2* xj = reinterpret_cast<const double2*>(Wc + 36 + 6 * j);
              const double2 b0 = xj[0], b1 = xj[1], b2 = xj[2];
              An[6 * lane + j] -= a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y;
            }
          }
        }
        __syncwarp();
        if (lane == 0) {
          const long long tc0 = dbg ? clock64() : 0;
          if (!chol6(An, 6, LI + static_cast<size_t>(c + 1) * 48)) s_ok = 0;
          if (dbg) { const long long tc1 = clock64(); s_dbg[1] += tc1 - tc0; s_dbg[0] += tc0 - s_mark; }
        }
      }
    }
    __syncthreads();
    HB_TICK(3);
  }
  // ---- corner (m x m, rhs carried as row m): right-looking Cholesky by the whole CTA, two barriers
  // per column (every thread recomputes the reciprocal square root of the pivot) ----
  for (int q = 0; q < m; ++q) {
    const double d = CC[static_cast<size_t>(q) * m + q];
    if (!(d > 0.0) && tid == 0) s_ok = 0;
    const double iv = rsqrt(d);
    __syncthreads();   // everyone has read the pivot before it is overwritten
    for (int r = q + tid; r <= m; r += kBandThreads) CC[static_cast<size_t>(r) * m + q] = (r == q) ? d * iv : CC[static_cast<size_t>(r) * m + q] * iv;
    __syncthreads();
    const int rem = m - q;  // rows q+1 .. m (row m = rhs), columns q+1 .. m-1
    for (int e = tid; e < rem * rem; e += kBandThreads) {
      const int u = q + 1 + e / rem, v = q + 1 + e % rem;
      if (v > u || v >= m) continue;
      CC[static_cast<size_t>(u) * m + v] -= CC[static_cast<size_t>(u) * m + q] * CC[static_cast<size_t>(v) * m + q];
    }
    // the next iteration's pivot read is ordered after these updates by its first barrier? no: it reads
    // CC[q+1][q+1] right away, so synchronise here
    __syncthreads();
  }
  // ---- the whole back substitution in warp 0 ----
  if (warp == 0) {
    HB_TICK(4);
    double* Xa = X + np;
    for (int q = m - 1; q >= 0; --q) {
      double xq = 0.0;
      if ((q & 31) == lane) { xq = CC[static_cast<size_t>(m) * m + q] / CC[static_cast<size_t>(q) * m + q]; Xa[q] = xq; }
      xq = __shfl_sync(0xffffffffu, xq, q & 31);
      for (int r = lane; r < q; r += 32) CC[static_cast<size_t>(m) * m + r] -= CC[static_cast<size_t>(q) * m + r] * xq;
      __syncwarp();
    }
    HB_TICK(5);
  }
  __syncthreads();
  // arrow contribution to every block's right-hand side, all at once: y_p -= AR^T x_a (row m of AR = y)
  {
    const double* Xa = X + np;
    for (int col = tid; col < np; col += kBandThreads) {
      double s = 0.0;
      for (int r = 0; r < m; ++r) s += AR[static_cast<size_t>(r) * np + col] * Xa[r];
      AR[static_cast<size_t>(m) * np + col] -= s;
    }
  }
  __syncthreads();
  if (warp == 0) {
    // block columns K-1 .. 0: lane = (component j = lane % 6, part p = lane / 6), 5 parts
    const int j = lane % 6, part = lane / 6;
    for (int c = K - 1; c >= 0; --c) {
      const double* Wc = W + static_cast<size_t>(c) * h * 6;
      const double* Li = LI + static_cast<size_t>(c) * 48 + 8;
      const int nb = min(h - 6, np - 6 * (c + 1));
      double s = 0.0, s2 = 0.0;
      if (part < 5) {   // band rows only; the arrow part was folded into y above
        int t = part;
        for (; t + 5 < nb; t += 10) {   // two independent accumulators
          s += Wc[static_cast<size_t>(6 + t) * 6 + j] * X[6 * (c + 1) + t];
          s2 += Wc[static_cast<size_t>(6 + t + 5) * 6 + j] * X[6 * (c + 1) + t + 5];
        }
        if (t < nb) s += Wc[static_cast<size_t>(6 + t) * 6 + j] * X[6 * (c + 1) + t];
        s += s2;
      }
      // sum the 5 parts of each component: lanes j, j+6, j+12, j+18, j+24
      double tot = s;
      tot += __shfl_sync(0xffffffffu, s, (lane + 6) & 31);
      tot += __shfl_sync(0xffffffffu, s, (lane + 12) & 31);
      tot += __shfl_sync(0xffffffffu, s, (lane + 18) & 31);
      tot += __shfl_sync(0xffffffffu, s, (lane + 24) & 31);
      // lanes 0..5 now hold the full sums (their partners are lanes j+6k < 30)
      double v = (lane < 6) ? AR[static_cast<size_t>(m) * np + 6 * c + lane] - tot : 0.0;
      // x = L^-T v: lane j sums Li[q][j] v_q over q >= j (six independent broadcasts)
      double xo = 0.0;
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const double vq = __shfl_sync(0xffffffffu, v, q);
        if (lane < 6 && q >= lane) xo += Li[q * 6 + lane] * vq;
      }
      if (lane < 6) X[6 * c + lane] = xo;
      __syncwarp();
    }
    HB_TICK(6);
  }
  __syncthreads();
  for (int e = tid; e < n; e += kBandThreads) x_out[e] = X[e];
  if (dbg && tid == 0) { t_acc[0] = s_dbg[0]; t_acc[7] = s_dbg[1]; for (int i = 0; i < 8; ++i) dbg[i] = t_acc[i]; }
  if (tid == 0) *spd_flag = s_ok;
}

}  // namespace hb
