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
// Two-sided elimination.  The factorisation is a chain of K dependent block-column steps, each bounded
// by FP64 dependent-issue latency (a 6x6 potf2 plus a panel solve), not by throughput.  The chain is cut
// in two: a separator of beta block columns in the middle decouples the columns above it from the
// columns below it, so chain 0 eliminates block columns 0 .. Kt-1 top-down while chain 1 eliminates
// K-1 .. Kt+beta bottom-up IN THE SAME STEPS (same barriers, disjoint warps, one look-ahead warp per
// chain on its own scheduler).  Chain 1 runs the very same code on the index-reversed matrix J P J
// (still block-banded) and accumulates its Schur updates of the separator and of the separator's
// arrow columns into a private, zero-initialised copy.  The copies are then merged into chain 0, which
// eliminates the beta separator columns, and the back substitution runs outwards from the separator
// in both directions at once.  Sequential depth: ~K/2 + beta steps instead of K.
// The arrow x arrow (corner) part of every step's trailing update is deferred: the corner receives
// C -= sum_c A_c A_c^T in one parallel pass after the last column, which takes 15 of 36 tiles out of
// every step.
//
// Per-chain workspace (doubles): W[ncol][h][6] band block-columns (h = 6 + 6 beta rows each: diagonal
//   block first; ncol = eliminated columns + separator columns), AR[m+1][6 ncol] arrow rows (+ the rhs as
//   last row), LI[ncol][48] reciprocal diagonals (6) and inverses (36, at +8) of the diagonal Cholesky
//   blocks, X[6 ncol].  Shared: CC[m+1][m] corner (+ rhs row), XA[m].
#pragma once
#include "hb200_solve.cuh"

namespace hb {

constexpr int kBandThreads = 512;

struct BandPlan { int Kt, Kb, bs; };   // columns eliminated by chain 0 / chain 1, separator columns
__host__ __device__ inline BandPlan band_plan(int K, int beta) {
  BandPlan p;
  if (K >= 2 * beta + 4) { p.bs = beta; p.Kt = (K - beta + 1) / 2; p.Kb = K - beta - p.Kt; }
  else { p.bs = 0; p.Kt = K; p.Kb = 0; }
  return p;
}

__host__ __device__ inline size_t band_workspace_doubles(int K, int beta, int m) {
  const BandPlan p = band_plan(K, beta);
  const size_t h = 6 + 6 * static_cast<size_t>(beta);
  const size_t n0 = p.Kt + p.bs, n1 = p.Kb ? p.Kb + p.bs : 0;
  size_t d = (n0 + n1) * h * 6;                  // W
  d += static_cast<size_t>(m + 1) * 6 * (n0 + n1);  // AR
  d += (n0 + n1) * 48;                           // LI
  d += 6 * (n0 + n1);                            // X
  d += static_cast<size_t>(m + 1) * m + m;       // CC, XA
  return d + 8;
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

// Clock read that waits for `dep` (BAR.SYNC is deferred-blocking: a bare clock read right after a barrier
// reports the barrier's issue time, not its release).
HB_DI long long clock_after(double dep) {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "d"(dep) : "memory");
  return t;
}

struct BandChain {
  double *W, *AR, *LI, *X;
  int Ke;     // block columns this chain eliminates in the two-sided phase
  int ncol;   // block columns it stores (Ke + separator)
  int npc;    // 6 * ncol
};

// Panel of block column s: solve x L^T = a for every row below the diagonal block (band rows, arrow rows,
// rhs row), one row per thread lt, lt += nthreads.
HB_DI void band_panel(const BandChain& C, int s, int h, int m, int lt, int nthreads) {
  double* Wc = C.W + static_cast<size_t>(s) * h * 6;
  const double* Lic = C.LI + static_cast<size_t>(s) * 48;
  const int nb = min(h - 6, C.npc - 6 * (s + 1));  // band rows below the diagonal block
  const int R = nb + m + 1;                          // + arrow rows + rhs row
  for (int t = lt; t < R; t += nthreads) {
    double* a = (t < nb) ? (Wc + static_cast<size_t>(6 + t) * 6) : (C.AR + static_cast<size_t>(t - nb) * C.npc + 6 * s);
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
}

// Trailing update of block column s in 6x6 tiles (band x band lower triangle, then arrow x band; the
// arrow x arrow part is deferred).  Worker = (tile, row i of the tile): 6 dots of length 6.  With `look`
// tile 0 (the next diagonal block) is left to the look-ahead warp.
HB_DI void band_update(const BandChain& C, int s, int h, int m, bool look, int worker, int nworkers) {
  double* Wc = C.W + static_cast<size_t>(s) * h * 6;
  const int nb = min(h - 6, C.npc - 6 * (s + 1));
  const int nbk = nb / 6;
  const int ng = (m + 1 + 5) / 6;
  const int ntri = nbk * (nbk + 1) / 2;
  const int ntiles = ntri + ng * nbk;
  for (int t = (look ? 6 : 0) + worker; t < ntiles * 6; t += nworkers) {
    const int tile = t / 6, i = t - 6 * tile;
    int gu, gv;
    if (tile < ntri) {
      gu = static_cast<int>((sqrtf(8.0f * tile + 1.0f) - 1.0f) * 0.5f);
      while (gu * (gu + 1) / 2 > tile) --gu;
      while ((gu + 1) * (gu + 2) / 2 <= tile) ++gu;
      gv = tile - gu * (gu + 1) / 2;
    } else {
      const int ta = tile - ntri;
      gu = nbk + ta / nbk;
      gv = ta - (ta / nbk) * nbk;
    }
    const bool ub = gu < nbk;
    const int ru = ub ? 0 : 6 * (gu - nbk) + i;   // arrow row index of u
    if (!ub && ru > m) continue;
    const double* xu = ub ? (Wc + static_cast<size_t>(6 + 6 * gu + i) * 6) : (C.AR + static_cast<size_t>(ru) * C.npc + 6 * s);
    double a[6], sd[6];
    {
      const double2* x2 = reinterpret_cast<const double2*>(xu);   // rows are 48 B: 16-byte aligned
      const double2 a0 = x2[0], a1 = x2[1], a2 = x2[2];
      a[0] = a0.x; a[1] = a0.y; a[2] = a1.x; a[3] = a1.y; a[4] = a2.x; a[5] = a2.y;
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const double2* v2 = reinterpret_cast<const double2*>(Wc + static_cast<size_t>(6 + 6 * gv + j) * 6);
      const double2 b0 = v2[0], b1 = v2[1], b2 = v2[2];
      sd[j] = a[0] * b0.x + a[1] * b0.y + a[2] * b1.x + a[3] * b1.y + a[4] * b2.x + a[5] * b2.y;
    }
    double* tgt = ub ? (C.W + (static_cast<size_t>(s + 1 + gv) * h + (6 * (gu - gv) + i)) * 6)   // band x band
                     : (C.AR + static_cast<size_t>(ru) * C.npc + 6 * (s + 1 + gv));              // arrow x band
#pragma unroll
    for (int j = 0; j < 6; ++j) tgt[j] -= sd[j];
  }
}

// Inverse of block column s's diagonal factor, for the back substitution (one thread per block column, all
// columns at once after the factorisation -- it used to sit in the per-step loop, where its long
// store->load dependent chain, not the look-ahead potf2, bounded the step):
// Li[i][j] = -(sum_{p=j}^{i-1} L[i][p] Li[p][j]) / L[i][i],  Li[j][j] = 1 / L[j][j]
HB_DI void band_block_inverse(const BandChain& C, int s, int h) {
  const double* Wc = C.W + static_cast<size_t>(s) * h * 6;
  const double* Lic = C.LI + static_cast<size_t>(s) * 48;
  double* Lo = C.LI + static_cast<size_t>(s) * 48 + 8;
  double L[36], Li[36], rd[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    rd[i] = Lic[i];
#pragma unroll
    for (int j = 0; j < 6; ++j) { L[6 * i + j] = (j < i) ? Wc[i * 6 + j] : 0.0; Li[6 * i + j] = 0.0; }
  }
#pragma unroll
  for (int jj = 0; jj < 6; ++jj) {
    Li[jj * 6 + jj] = rd[jj];
#pragma unroll
    for (int ii = jj + 1; ii < 6; ++ii) {
      double acc = 0.0;
#pragma unroll
      for (int pp = jj; pp < ii; ++pp) acc -= L[ii * 6 + pp] * Li[pp * 6 + jj];
      Li[ii * 6 + jj] = acc * rd[ii];
    }
  }
#pragma unroll
  for (int e = 0; e < 36; ++e) Lo[e] = Li[e];
}

// Look-ahead (one warp): the next diagonal block receives row i of its update A -= X X^T from this step's
// panel rows (tile 0) in lanes 0..5, then lane 0 factors it.
HB_DI bool band_lookahead(const BandChain& C, int s, int h, int lane) {
  const double* Wc = C.W + static_cast<size_t>(s) * h * 6;
  double* An = C.W + static_cast<size_t>(s + 1) * h * 6;
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
  bool ok = true;
  if (lane == 0) ok = chol6(An, 6, C.LI + static_cast<size_t>(s + 1) * 48);
  return ok;
}

// One block column of the back substitution (one warp): x_c = L_cc^-T (y_c - sum_band L_rc^T x_r); the arrow
// part was folded into y beforehand.  lane = (component j = lane % 6, part p = lane / 6), 5 parts.
HB_DI void band_backsub_column(const BandChain& C, int c, int h, int m, int lane) {
  const int j = lane % 6, part = lane / 6;
  const double* Wc = C.W + static_cast<size_t>(c) * h * 6;
  const double* Li = C.LI + static_cast<size_t>(c) * 48 + 8;
  const int nb = min(h - 6, C.npc - 6 * (c + 1));
  double sa = 0.0, s2 = 0.0;
  if (part < 5) {
    int t = part;
    for (; t + 5 < nb; t += 10) {   // two independent accumulators
      sa += Wc[static_cast<size_t>(6 + t) * 6 + j] * C.X[6 * (c + 1) + t];
      s2 += Wc[static_cast<size_t>(6 + t + 5) * 6 + j] * C.X[6 * (c + 1) + t + 5];
    }
    if (t < nb) sa += Wc[static_cast<size_t>(6 + t) * 6 + j] * C.X[6 * (c + 1) + t];
    sa += s2;
  }
  // sum the 5 parts of each component: lanes j, j+6, j+12, j+18, j+24
  double tot = sa;
  tot += __shfl_sync(0xffffffffu, sa, (lane + 6) & 31);
  tot += __shfl_sync(0xffffffffu, sa, (lane + 12) & 31);
  tot += __shfl_sync(0xffffffffu, sa, (lane + 18) & 31);
  tot += __shfl_sync(0xffffffffu, sa, (lane + 24) & 31);
  // lanes 0..5 now hold the full sums (their partners are lanes j+6k < 30)
  const double v = (lane < 6) ? C.AR[static_cast<size_t>(m) * C.npc + 6 * c + lane] - tot : 0.0;
  // x = L^-T v: lane j sums Li[q][j] v_q over q >= j (six independent broadcasts)
  double xo = 0.0;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    const double vq = __shfl_sync(0xffffffffu, v, q);
    if (lane < 6 && q >= lane) xo += Li[q * 6 + lane] * vq;
  }
  if (lane < 6) C.X[6 * c + lane] = xo;
  __syncwarp();
}

template <bool SMEM>
__global__ void __launch_bounds__(kBandThreads) band_solve_kernel(const double* __restrict__ sys, int n, int K, int beta,
                                                                  double* __restrict__ ws_global, double* __restrict__ x_out,
                                                                  int* __restrict__ spd_flag, long long* __restrict__ dbg) {
  extern __shared__ double s_band[];
  double* ws = SMEM ? s_band : ws_global;
  const int np = 6 * K, m = n - np, h = 6 + 6 * beta, h6 = h * 6;
  const BandPlan pl = band_plan(K, beta);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // Roles.  Each chain owns one scheduler for its latency-critical work (the arbiter favours the highest
  // warp id of a scheduler): scheduler 0 = chain 0 (warp 12 look-ahead, warp 8 lane 0 block inverses),
  // scheduler 1 = chain 1 (warp 13, warp 9 lane 0); their other warps idle during the update phase.
  // The warps of schedulers 2 and 3 do the trailing updates: 2,3,6,7 -> chain 0, 10,11,14,15 -> chain 1.
  const int sched = warp & 3;
  const int my_chain = (sched < 2) ? sched : (warp >> 3);
  BandChain C0, C1;
  {
    double* p = ws;
    BandChain* cc[2] = {&C0, &C1};
    for (int i = 0; i < 2; ++i) {
      BandChain& C = *cc[i];
      C.Ke = i ? pl.Kb : pl.Kt;
      C.ncol = (i && pl.Kb == 0) ? 0 : C.Ke + pl.bs;
      C.npc = 6 * C.ncol;
      C.W = p; p += static_cast<size_t>(C.ncol) * h6;
      C.AR = p; p += static_cast<size_t>(m + 1) * C.npc;
      C.LI = p; p += static_cast<size_t>(C.ncol) * 48;
      C.X = p; p += C.npc;
    }
    ws = p;
  }
  double* CC = ws;
  double* XA = CC + static_cast<size_t>(m + 1) * m;
  const BandChain C = my_chain ? C1 : C0;           // this thread's chain in the update phase (registers)
  const BandChain CP = (tid >> 8) ? C1 : C0;        // ... and in the panel phase (threads 0..255 / 256..511)
  const double* S = sys;
  const double* b = sys + static_cast<size_t>(n) * n;
  __shared__ int s_ok;
  if (tid == 0) s_ok = 1;
  // ---- gather: chain 0 reads P top-down, chain 1 reads J P J (index reversal) and starts its copy of the
  // separator and of the separator's arrow columns at zero (they only accumulate updates) ----
  {
    const int N1 = C1.npc;
    for (int e = tid; e < C0.ncol * h6; e += kBandThreads) {
      const int c = e / h6, rem = e - c * h6;
      const int i = rem / 6, j = rem - 6 * i;
      const int row = 6 * c + i, col = 6 * c + j;
      C0.W[e] = (row < C0.npc && row >= col) ? S[static_cast<size_t>(row) * n + col] : 0.0;
    }
    for (int e = tid; e < C1.ncol * h6; e += kBandThreads) {
      const int c = e / h6, rem = e - c * h6;
      const int i = rem / 6, j = rem - 6 * i;
      const int rr = 6 * c + i, rc = 6 * c + j;          // reversed (chain-local) row / column
      double v = 0.0;
      if (c < C1.Ke && rr < N1 && rr >= rc) v = S[static_cast<size_t>(np - 1 - rc) * n + (np - 1 - rr)];
      C1.W[e] = v;
    }
    for (int e = tid; e < (m + 1) * C0.npc; e += kBandThreads) {
      const int r = e / C0.npc, col = e - r * C0.npc;
      C0.AR[e] = (r < m) ? S[static_cast<size_t>(np + r) * n + col] : b[col];
    }
    for (int e = tid; e < (m + 1) * N1; e += kBandThreads) {
      const int r = e / N1, rc = e - r * N1;
      const int col = np - 1 - rc;
      C1.AR[e] = (rc < 6 * C1.Ke) ? ((r < m) ? S[static_cast<size_t>(np + r) * n + col] : b[col]) : 0.0;
    }
    for (int e = tid; e < (m + 1) * m; e += kBandThreads) {
      const int r = e / m, q = e - r * m;
      CC[e] = (r < m) ? S[static_cast<size_t>(np + r) * n + np + q] : b[np + q];
    }
  }
  __syncthreads();
  long long t_mark = clock64(), t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define HB_TICK(i) do { if (dbg && tid == 0) { const long long now = clock64(); t_acc[i] += now - t_mark; t_mark = now; } } while (0)
  // ---- two-sided factorisation + forward substitution: step s eliminates block column s of BOTH chains.
  // The 6x6 Cholesky of a chain's NEXT diagonal block is done by its look-ahead warp inside the
  // trailing-update phase, as soon as that block has received its update: off the critical path. ----
  if (tid == 0 && C0.Ke > 0) { if (!chol6(C0.W, 6, C0.LI)) s_ok = 0; }
  if (tid == 32 && C1.Ke > 0) { if (!chol6(C1.W, 6, C1.LI)) s_ok = 0; }
  __syncthreads();
  HB_TICK(0);
  const int steps = max(C0.Ke, C1.Ke);
  const int urank = ((warp >> 2) & 1) * 2 + (sched - 2);   // update warps: 0..3 within their chain
  __shared__ long long s_ts[8][8];
  for (int s = 0; s < steps; ++s) {
    const bool rec = dbg && s >= 4 && s < 12;
    if (rec && tid == 0) s_ts[s - 4][0] = clock_after(C0.LI[static_cast<size_t>(s) * 48]);      // panel may start (chol(s) visible)
    if (s < CP.Ke) band_panel(CP, s, h, m, tid & 255, 256);
    if (rec && tid == 0) s_ts[s - 4][1] = clock_after(C0.W[static_cast<size_t>(s) * h6 + 36]);  // own panel row written
    __syncthreads();
    if (rec && tid == 2 * 32) s_ts[s - 4][2] = clock_after(C0.W[static_cast<size_t>(s) * h6 + 36]);   // update warp released
    if (rec && tid == 12 * 32) s_ts[s - 4][4] = clock_after(C0.W[static_cast<size_t>(s) * h6 + 36]);  // look-ahead warp released
    if (s < C.Ke) {
      const bool look = s + 1 < C.Ke;
      if (sched >= 2) band_update(C, s, h, m, look, urank * 32 + lane, 4 * 32);
      else if ((warp >> 2) == 3 && look) { if (!band_lookahead(C, s, h, lane)) s_ok = 0; }
    }
    if (rec && tid == 2 * 32) s_ts[s - 4][3] = clock_after(C0.W[static_cast<size_t>(s + 1) * h6 + 6 * 6]);   // update warp done (approx)
    if (rec && tid == 12 * 32) s_ts[s - 4][5] = clock_after(C0.LI[static_cast<size_t>(s + 1) * 48]);        // chol(s+1) done
    __syncthreads();
  }
  if (dbg && tid < 64) dbg[8 + tid] = s_ts[tid >> 3][tid & 7];
  // ---- merge chain 1's copy of the separator (index-reversed) and of its arrow columns into chain 0,
  // then chain 0 eliminates the separator columns with all 8 update warps ----
  if (pl.Kb) {
    const int nsep = 6 * pl.bs, N1 = C1.npc;
    for (int e = tid; e < nsep * nsep; e += kBandThreads) {
      const int u = e / nsep, v = e - u * nsep;
      if (v > u) continue;
      const int ru = N1 - 1 - u, rv = N1 - 1 - v;   // ru <= rv: stored in block column ru / 6
      C0.W[(static_cast<size_t>(C0.Ke + v / 6) * h + (u - 6 * (v / 6))) * 6 + v % 6] +=
          C1.W[(static_cast<size_t>(ru / 6) * h + (rv - 6 * (ru / 6))) * 6 + ru % 6];
    }
    for (int e = tid; e < (m + 1) * nsep; e += kBandThreads) {
      const int r = e / nsep, v = e - r * nsep;
      C0.AR[static_cast<size_t>(r) * C0.npc + 6 * C0.Ke + v] += C1.AR[static_cast<size_t>(r) * N1 + (N1 - 1 - v)];
    }
    __syncthreads();
    if (tid == 0 && !chol6(C0.W + static_cast<size_t>(C0.Ke) * h6, 6, C0.LI + static_cast<size_t>(C0.Ke) * 48)) s_ok = 0;
    __syncthreads();
    const int urank8 = (warp >> 2) * 2 + (sched - 2);
    for (int s = C0.Ke; s < C0.ncol; ++s) {
      band_panel(C0, s, h, m, tid, kBandThreads);
      __syncthreads();
      const bool look = s + 1 < C0.ncol;
      if (sched >= 2) band_update(C0, s, h, m, look, urank8 * 32 + lane, 8 * 32);
      else if (warp == 12 && look) { if (!band_lookahead(C0, s, h, lane)) s_ok = 0; }
      __syncthreads();
    }
  }
  HB_TICK(2);
  // ---- block inverses of every eliminated column (one thread each), and the deferred corner update
  // C -= sum over all eliminated columns of (arrow panel)(arrow panel)^T, rhs row included (row m):
  // one (u, v) pair per warp pass, lanes stride the columns (conflict-free), butterfly reduction ----
  for (int c = tid; c < C0.ncol + C1.Ke; c += kBandThreads) {
    if (c < C0.ncol) band_block_inverse(C0, c, h);
    else band_block_inverse(C1, c - C0.ncol, h);
  }
  {
    const int n0 = C0.npc, n1 = 6 * C1.Ke;
    const int npairs = (m + 1) * m;   // (u, v) with v <= u are used
    for (int e = warp; e < npairs; e += kBandThreads / 32) {
      const int u = e / m, v = e - u * m;
      if (v > u) continue;
      const double* au = C0.AR + static_cast<size_t>(u) * C0.npc;
      const double* av = C0.AR + static_cast<size_t>(v) * C0.npc;
      double a0 = 0.0, a1 = 0.0;
      int col = lane;
      for (; col + 32 < n0; col += 64) { a0 += au[col] * av[col]; a1 += au[col + 32] * av[col + 32]; }
      if (col < n0) a0 += au[col] * av[col];
      au = C1.AR + static_cast<size_t>(u) * C1.npc;
      av = C1.AR + static_cast<size_t>(v) * C1.npc;
      for (col = lane; col + 32 < n1; col += 64) { a0 += au[col] * av[col]; a1 += au[col + 32] * av[col + 32]; }
      if (col < n1) a0 += au[col] * av[col];
      a0 += a1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      if (lane == 0) CC[e] -= a0;
    }
  }
  __syncthreads();
  HB_TICK(3);
  // ---- corner (m x m, rhs carried as row m): right-looking Cholesky by the whole CTA ----
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
    __syncthreads();
  }
  HB_TICK(4);
  if (warp == 0) {   // back substitution of the corner
    for (int q = m - 1; q >= 0; --q) {
      double xq = 0.0;
      if ((q & 31) == lane) { xq = CC[static_cast<size_t>(m) * m + q] / CC[static_cast<size_t>(q) * m + q]; XA[q] = xq; }
      xq = __shfl_sync(0xffffffffu, xq, q & 31);
      for (int r = lane; r < q; r += 32) CC[static_cast<size_t>(m) * m + r] -= CC[static_cast<size_t>(q) * m + r] * xq;
      __syncwarp();
    }
  }
  __syncthreads();
  HB_TICK(5);
  // arrow contribution to every eliminated block's right-hand side, all at once: y_p -= AR^T x_a (row m of AR = y)
  {
    const int n0 = C0.npc, n1 = 6 * C1.Ke;
    for (int e = tid; e < n0 + n1; e += kBandThreads) {
      const BandChain& Q = e < n0 ? C0 : C1;
      const int col = e < n0 ? e : e - n0;
      double sacc = 0.0;
      for (int r = 0; r < m; ++r) sacc += Q.AR[static_cast<size_t>(r) * Q.npc + col] * XA[r];
      Q.AR[static_cast<size_t>(m) * Q.npc + col] -= sacc;
    }
  }
  __syncthreads();
  // separator columns first (chain 0, warp 0), then outwards: chain 0 in warp 0, chain 1 in warp 1
  if (warp == 0) {
    for (int c = C0.ncol - 1; c >= C0.Ke; --c) band_backsub_column(C0, c, h, m, lane);
  }
  __syncthreads();
  if (pl.Kb) {
    for (int v = tid; v < 6 * pl.bs; v += kBandThreads) C1.X[C1.npc - 1 - v] = C0.X[6 * C0.Ke + v];
    __syncthreads();
  }
  if (warp == 0) {
    for (int c = C0.Ke - 1; c >= 0; --c) band_backsub_column(C0, c, h, m, lane);
  } else if (warp == 1) {
    for (int c = C1.Ke - 1; c >= 0; --c) band_backsub_column(C1, c, h, m, lane);
  }
  __syncthreads();
  HB_TICK(6);
  for (int e = tid; e < C0.npc; e += kBandThreads) x_out[e] = C0.X[e];
  for (int e = tid; e < 6 * C1.Ke; e += kBandThreads) x_out[np - 1 - e] = C1.X[e];
  for (int e = tid; e < m; e += kBandThreads) x_out[np + e] = XA[e];
  if (dbg && tid == 0) { for (int i = 0; i < 8; ++i) dbg[i] = t_acc[i]; }
  if (tid == 0) *spd_flag = s_ok;
}

}  // namespace hb
