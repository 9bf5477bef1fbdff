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
//   AR[m+1][np] arrow rows (+ the rhs as last row), CC[m+1][m] corner (+ rhs row), LI[K][8] reciprocal
//   diagonals of the diagonal Cholesky blocks, X[n].
#pragma once
#include "hb200_solve.cuh"

namespace hb {

constexpr int kBandThreads = 512;

__host__ __device__ inline size_t band_workspace_doubles(int K, int beta, int m) {
  const size_t h = 6 + 6 * static_cast<size_t>(beta);
  const size_t np = 6 * static_cast<size_t>(K);
  return static_cast<size_t>(K) * h * 6 + (m + 1) * np + static_cast<size_t>(m + 1) * m + static_cast<size_t>(K) * 8 + np + m;
}

// Right-looking Cholesky of a 6x6 SPD block (lower, row-major with stride ld), division-free:
// inv[j] = 1 / L[j][j] comes straight out of rsqrt.
HB_DI bool chol6(double* A, int ld, double* inv /*6*/) {
  double L[21];  // packed lower: L[i(i+1)/2 + j]
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) L[i * (i + 1) / 2 + j] = A[i * ld + j];
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

__global__ void __launch_bounds__(kBandThreads) band_solve_kernel(const double* __restrict__ sys, int n, int K, int beta, int use_smem,
                                                                  double* __restrict__ ws_global, double* __restrict__ x_out,
                                                                  int* __restrict__ spd_flag) {
  extern __shared__ double s_band[];
  double* ws = use_smem ? s_band : ws_global;
  const int np = 6 * K, m = n - np, h = 6 + 6 * beta;
  double* W = ws;
  double* AR = W + static_cast<size_t>(K) * h * 6;
  double* CC = AR + static_cast<size_t>(m + 1) * np;
  double* LI = CC + static_cast<size_t>(m + 1) * m;
  double* X = LI + static_cast<size_t>(K) * 8;
  const double* S = sys;
  const double* b = sys + static_cast<size_t>(n) * n;
  const int tid = threadIdx.x;
  __shared__ int s_ok;
  if (tid == 0) s_ok = 1;
  // ---- gather the band, the arrow and the corner from the dense system ----
  for (int e = tid; e < K * h * 6; e += kBandThreads) {
    const int c = e / (h * 6), rem = e - c * h * 6;
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
  __syncthreads();
  // ---- factorisation + forward substitution, one control-point block column per step ----
  for (int c = 0; c < K; ++c) {
    double* Wc = W + static_cast<size_t>(c) * h * 6;
    double* Lic = LI + static_cast<size_t>(c) * 8;
    if (tid == 0) {
      if (!chol6(Wc, 6, Lic)) s_ok = 0;
    }
    __syncthreads();
    const int nb = min(h - 6, np - 6 * (c + 1));  // band rows below the diagonal block
    const int R = nb + m + 1;                      // + arrow rows + rhs row
    // panel: X_row = A_row * L^-T  (6 values per row)
    for (int t = tid; t < R; t += kBandThreads) {
      double* a = (t < nb) ? (Wc + static_cast<size_t>(6 + t) * 6) : (AR + static_cast<size_t>(t - nb) * np + 6 * c);
      double v[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) v[q] = a[q];
#pragma unroll
      for (int j = 0; j < 6; ++j) {  // x L^T = a  (right-looking forward substitution)
        v[j] *= Lic[j];
#pragma unroll
        for (int q = j + 1; q < 6; ++q) v[q] -= v[j] * Wc[q * 6 + j];
      }
#pragma unroll
      for (int q = 0; q < 6; ++q) a[q] = v[q];
    }
    __syncthreads();
    // trailing update in 6x6 tiles: groups = band blocks below the diagonal, then arrow rows in sixes
    // (the rhs row is the last arrow row).  Thread = (tile, row i of the tile): 6 dots of length 6.
    {
      const int nbk = nb / 6;
      const int ng = (m + 1 + 5) / 6;
      const int G = nbk + ng;
      const int ntiles = G * (G + 1) / 2;
      for (int t = tid; t < ntiles * 6; t += kBandThreads) {
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
#pragma unroll
        for (int q = 0; q < 6; ++q) a[q] = xu[q];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const int rv = vb ? 0 : 6 * (gv - nbk) + j;
          const double* xv = vb ? (Wc + static_cast<size_t>(6 + 6 * gv + j) * 6) : (AR + static_cast<size_t>(min(rv, m)) * np + 6 * c);
          double acc = 0.0;
#pragma unroll
          for (int q = 0; q < 6; ++q) acc += a[q] * xv[q];
          sd[j] = acc;
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
    }
    __syncthreads();
  }
  // ---- corner: dense Cholesky of CC (m x m) with the rhs row carried along ----
  for (int q = 0; q < m; ++q) {
    if (tid == 0) {
      const double d = CC[static_cast<size_t>(q) * m + q];
      if (!(d > 0.0)) s_ok = 0;
      CC[static_cast<size_t>(q) * m + q] = sqrt(d);
    }
    __syncthreads();
    const double dq = CC[static_cast<size_t>(q) * m + q];
    for (int r = q + 1 + tid; r <= m; r += kBandThreads) CC[static_cast<size_t>(r) * m + q] /= dq;
    __syncthreads();
    const int rem = m - q;  // rows q+1..m
    for (int e = tid; e < rem * rem; e += kBandThreads) {
      const int u = q + 1 + e / rem, v = q + 1 + e % rem;
      if (v > u || v >= m) continue;
      CC[static_cast<size_t>(u) * m + v] -= CC[static_cast<size_t>(u) * m + q] * CC[static_cast<size_t>(v) * m + q];
    }
    __syncthreads();
  }
  // ---- back substitution: corner (warp 0), then block columns K-1 .. 0 ----
  double* Xa = X + np;
  if (tid < 32) {
    // y = row m of CC; lanes own entries r = lane, lane+32, ...
    for (int q = m - 1; q >= 0; --q) {
      double xq = 0.0;
      if ((q & 31) == tid) { xq = CC[static_cast<size_t>(m) * m + q] / CC[static_cast<size_t>(q) * m + q]; Xa[q] = xq; }
      xq = __shfl_sync(0xffffffffu, xq, q & 31);
      for (int r = tid; r < q; r += 32) CC[static_cast<size_t>(m) * m + r] -= CC[static_cast<size_t>(q) * m + r] * xq;
      __syncwarp();
    }
  }
  __syncthreads();
  __shared__ double s_rhs[6];
  const int warp = tid >> 5, lane = tid & 31;
  for (int c = K - 1; c >= 0; --c) {
    const double* Wc = W + static_cast<size_t>(c) * h * 6;
    const int nb = min(h - 6, np - 6 * (c + 1));
    if (warp < 6) {
      const int j = warp;
      double s = 0.0;
      for (int t = lane; t < nb + m; t += 32) {
        if (t < nb) s += Wc[static_cast<size_t>(6 + t) * 6 + j] * X[6 * (c + 1) + t];
        else s += AR[static_cast<size_t>(t - nb) * np + 6 * c + j] * Xa[t - nb];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) s_rhs[j] = AR[static_cast<size_t>(m) * np + 6 * c + j] - s;
    }
    __syncthreads();
    if (tid == 0) {
      const double* Lic = LI + static_cast<size_t>(c) * 8;
      double v[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) v[q] = s_rhs[q];
#pragma unroll
      for (int j = 5; j >= 0; --j) {  // L^T x = rhs
        v[j] *= Lic[j];
#pragma unroll
        for (int q = 0; q < j; ++q) v[q] -= v[j] * Wc[j * 6 + q];
      }
#pragma unroll
      for (int q = 0; q < 6; ++q) X[6 * c + q] = v[q];
    }
    __syncthreads();
  }
  for (int e = tid; e < n; e += kBandThreads) x_out[e] = X[e];
  if (tid == 0) *spd_flag = s_ok;
}

}  // namespace hb
