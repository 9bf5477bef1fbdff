// Multi-CTA solve of the reduced system for LONG windows (sm_100a): block cyclic reduction.
//
// band_solve_kernel factors the block-banded + arrowhead system (SysLayout) in ONE CTA: K/2 + beta dependent
// block-column steps, 0.6 ms at K = 200 and 1.4 ms at K = 500 -- more than every other kernel of the iteration
// together, and replicated on every rank of a multi-GPU solve.  Here the chain is replaced by a tree.
// The pose part is cut into super-blocks of beta control points (nb = 6 beta dofs): a half-bandwidth of beta
// blocks makes the super-block matrix block-TRIDIAGONAL (D_j on the diagonal, B_j = A[j+1][j] below it), with the
// arrow rows F_j (m x nb) and the corner C (m x m) attached.  Cyclic reduction eliminates every other
// super-block of the current level at once -- they do not couple to each other -- one CTA per eliminated node:
//
//   phase 1 (odd node i, neighbours a = i - s, c = i + s, s = 2^level):
//       L L^T = D_i ;  [Y_a | Y_c | Y_f | y] = L^-1 [B_a | B_i^T | F_i^T | b_i]          (kept for the way back)
//   phase 2 (even node e): Schur complements of its two eliminated neighbours
//       D_e -= Y_c(e-s)^T Y_c(e-s) + Y_a(e+s)^T Y_a(e+s)        F_e, b_e alike
//       B_e  = -Y_c(e+s)^T Y_a(e+s)                             (new coupling e -> e + 2s)
//
// and the even nodes form the next, half as long, block-tridiagonal level: ceil(log2(K / beta)) levels instead
// of K / 2 steps.  The arrow x arrow corner receives C -= sum_i Y_f(i)^T Y_f(i) once, after the last level
// (per-CTA partial sums reduced in a fixed order), is factored by CTA 0, and the solution flows back down the
// tree: x_i = L^-T (y - Y_a x_a - Y_c x_c - Y_f x_arrow), one level per grid synchronisation.
// LM damping mu clamp(diag H) and the constant-dof mask are applied while gathering, as in band_solve_kernel.
// Cooperative launch; everything between the levels lives in a global-memory workspace (a few MB, L2-resident).
// The reference hands this system to CHOLMOD through SPARSE_NORMAL_CHOLESKY, single-threaded
// (reference internal/hyper/optimizers/ceres/optimizer.cpp:41,46-48).
#pragma once
#include "hb200_band.cuh"

namespace hb {

constexpr int kBcrThreads = 512;
constexpr int kBcrMaxNb = 48;     // 6 * beta <= 48: landmark tracks of up to 8 control points

struct BcrPlan {
  int nb, nsb, m, levels;
  long long oD, oB, oF, ob, oL, oYa, oYc, oYf, oy, node_stride;   // per-node offsets (doubles)
  long long oCC, oCCp, oXa, total;                                 // global part (after nsb nodes)
};
__host__ __device__ inline BcrPlan bcr_plan(int K, int beta, int m, int num_ctas) {
  BcrPlan p;
  p.nb = 6 * beta; p.nsb = (K + beta - 1) / beta; p.m = m;
  p.levels = 0;
  while ((1 << p.levels) < p.nsb) p.levels += 1;
  const long long nn = static_cast<long long>(p.nb) * p.nb, mn = static_cast<long long>(m) * p.nb;
  p.oD = 0; p.oB = nn; p.oF = 2 * nn; p.ob = p.oF + mn; p.oL = p.ob + p.nb; p.oYa = p.oL + nn; p.oYc = p.oYa + nn;
  p.oYf = p.oYc + nn; p.oy = p.oYf + mn; p.node_stride = (p.oy + p.nb + 1) & ~1LL;
  p.oCC = p.node_stride * p.nsb;
  p.oCCp = p.oCC + static_cast<long long>(m + 1) * m;
  p.oXa = p.oCCp + static_cast<long long>(num_ctas) * (m + 1) * m;
  p.total = p.oXa + m + 2;
  return p;
}

// software grid barrier (all CTAs co-resident: cooperative launch).  Monotonic counter, one arrival per CTA.
HB_DI void bcr_grid_sync(unsigned int* counter, unsigned int* phase_local) {
  __syncthreads();
  if (threadIdx.x == 0) {
    *phase_local += 1;
    const unsigned int target = *phase_local * gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    while (*reinterpret_cast<volatile unsigned int*>(counter) < target) {}
    __threadfence();
  }
  __syncthreads();
}

// In-place Cholesky of the nb x nb SPD matrix in shared memory (row stride ld) by the whole CTA: fixed element
// ownership, two barriers per column, reciprocal diagonal to inv[].  Returns false on a non-positive pivot.
HB_DI bool bcr_chol(double* A, int ld, int nb, double* inv, int* s_flag) {
  const int tid = threadIdx.x;
  const int ntri = nb * (nb + 1) / 2;
  // element ownership decoded once (nb <= 48: at most 3 elements per thread)
  int eu[3], ev[3];
#pragma unroll
  for (int w = 0; w < 3; ++w) {
    const int e = tid + w * kBcrThreads;
    eu[w] = 0; ev[w] = -1;
    if (e < ntri) {
      int u = static_cast<int>((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
      while (u * (u + 1) / 2 > e) --u;
      while ((u + 1) * (u + 2) / 2 <= e) ++u;
      eu[w] = u; ev[w] = e - u * (u + 1) / 2;
    }
  }
  for (int q = 0; q < nb; ++q) {
    // the pivot A[q][q] is not written inside the loop (its square root goes to inv[]): no barrier before reading it
    const double d = A[q * ld + q];
    const double iv = rsqrt(d);
    if (tid == 0) { if (!(d > 0.0)) *s_flag = 0; inv[q] = iv; }
    for (int r = q + 1 + tid; r < nb; r += kBcrThreads) A[r * ld + q] *= iv;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 3; ++w)
      if (ev[w] > q) A[eu[w] * ld + ev[w]] -= A[eu[w] * ld + q] * A[ev[w] * ld + q];
    __syncthreads();
  }
  for (int q = tid; q < nb; q += kBcrThreads) A[q * ld + q] = 1.0 / inv[q];   // L_qq, consistent with the reciprocal used by the solves
  __syncthreads();
  return true;
}

__global__ void __launch_bounds__(kBcrThreads) bcr_solve_kernel(const double* __restrict__ sys, SysLayout lay, BcrPlan pl, double* ws,
                                                                unsigned int* barrier, double* x_out, int* __restrict__ spd_flag,
                                                                const SolverState* __restrict__ st, const unsigned char* __restrict__ fixed,
                                                                double* __restrict__ Dout) {
  extern __shared__ double s_bcr[];
  __shared__ int s_ok;
  __shared__ unsigned int s_phase;
  const int tid = threadIdx.x, cta = blockIdx.x, nctas = gridDim.x;
  const int nb = pl.nb, nsb = pl.nsb, m = pl.m, np = lay.np, n = lay.n, beta = lay.beta;
  if (tid == 0) { s_ok = 1; s_phase = *reinterpret_cast<volatile unsigned int*>(barrier) / gridDim.x; }   // counter persists across launches (multiple of gridDim.x)
  __syncthreads();
  auto node = [&](int j) -> double* { return ws + pl.node_stride * j; };
  const double mu = 1.0 / st->radius;
  // ---- gather: band-only packed system -> block-tridiagonal nodes, damping + mask applied ----
  for (int j = cta; j < nsb; j += nctas) {
    double* N = node(j);
    const int r0 = nb * j;   // first dof of the super-block
    for (int e = tid; e < nb * nb; e += kBcrThreads) {
      const int u = e / nb, v = e - u * nb;
      const int row = r0 + u, col = r0 + v;
      double d = 0.0;
      if (row < np && col < np) {
        const int hi = max(row, col), lo = min(row, col);
        d = sys[sys_index(lay, hi, lo)];
        if (row == col) d += mu * fmin(fmax(sys[lay.oD + row], 1e-6), 1e32);
        if (fixed[row] | fixed[col]) d = (row == col) ? 1.0 : 0.0;
      } else if (row == col) d = 1.0;   // padding of the last super-block
      N[pl.oD + e] = d;
      // B_j = A[j+1][j]: rows of super-block j + 1, columns of j; inside the band iff row - 6 (col / 6) < h
      const int rowb = r0 + nb + u;
      double b = 0.0;
      if (rowb < np && col < np && rowb - 6 * (col / 6) < lay.h && !(fixed[rowb] | fixed[col])) b = sys[sys_index(lay, rowb, col)];
      N[pl.oB + e] = b;
    }
    for (int e = tid; e < m * nb; e += kBcrThreads) {
      const int r = e / nb, v = e - r * nb;
      const int col = r0 + v;
      double f = 0.0;
      if (col < np && !(fixed[np + r] | fixed[col])) f = sys[lay.oA + static_cast<long long>(r) * np + col];
      N[pl.oF + e] = f;
    }
    for (int v = tid; v < nb; v += kBcrThreads) {
      const int col = r0 + v;
      N[pl.ob + v] = (col < np && !fixed[col]) ? sys[lay.ob + col] : 0.0;
    }
  }
  if (cta == 0) {
    double* CC = ws + pl.oCC;
    for (int e = tid; e < (m + 1) * m; e += kBcrThreads) {
      const int r = e / m, q = e - r * m;
      double v;
      if (r < m) {
        v = (q <= r) ? sys[sys_index(lay, np + r, np + q)] : 0.0;
        if (q == r) v += mu * fmin(fmax(sys[lay.oD + np + r], 1e-6), 1e32);
        if (fixed[np + r] | fixed[np + q]) v = (q == r) ? 1.0 : 0.0;
      } else v = fixed[np + q] ? 0.0 : sys[lay.ob + np + q];
      CC[e] = v;
    }
    for (int a = tid; a < n; a += kBcrThreads) Dout[a] = fmin(fmax(sys[lay.oD + a], 1e-6), 1e32);
    if (tid == 0) *spd_flag = 1;   // (cleared at the end by any CTA that met a non-positive pivot)
  }
  bcr_grid_sync(barrier, &s_phase);

  // shared-memory carve-up (doubles): L [nb][ldl] | inv [nb] | R [nb][ldr] (phase 1) ; staged Y matrices (phase 2)
  const int ldl = nb | 1;
  const int ncols = 2 * nb + m + 1, ldr = ncols | 1;
  double* sL = s_bcr;
  double* sInv = sL + nb * ldl;
  double* sR = sInv + nb + (nb & 1);
  // ---- reduction levels ----
  for (int lv = 0; lv < pl.levels; ++lv) {
    const int s = 1 << lv;
    // phase 1: odd nodes i = s, 3s, 5s, ...
    for (int i = s + 2 * s * cta; i < nsb; i += 2 * s * nctas) {
      double* N = node(i);
      const int a = i - s, c = i + s;
      const bool has_c = c < nsb;
      const double* Ba = node(a) + pl.oB;    // A[i][a]: rows i, columns a
      for (int e = tid; e < nb * nb; e += kBcrThreads) {
        const int u = e / nb, v = e - u * nb;
        sL[u * ldl + v] = N[pl.oD + e];
        sR[u * ldr + v] = Ba[e];                                       // column v of B_a, row u
        sR[u * ldr + nb + v] = has_c ? N[pl.oB + v * nb + u] : 0.0;   // (B_i^T)[u][v] = B_i[v][u]
      }
      for (int e = tid; e < m * nb; e += kBcrThreads) {
        const int r = e / nb, u = e - r * nb;
        sR[u * ldr + 2 * nb + r] = N[pl.oF + e];                       // (F_i^T)[u][r]
      }
      for (int u = tid; u < nb; u += kBcrThreads) sR[u * ldr + 2 * nb + m] = N[pl.ob + u];
      __syncthreads();
      bcr_chol(sL, ldl, nb, sInv, &s_ok);
      // forward substitution, one right-hand-side column per thread (L broadcast from shared memory)
      for (int t = tid; t < ncols; t += kBcrThreads) {
        for (int r = 0; r < nb; ++r) {
          double acc0 = sR[r * ldr + t], acc1 = 0.0;
          int q = 0;
          for (; q + 1 < r; q += 2) { acc0 -= sL[r * ldl + q] * sR[q * ldr + t]; acc1 -= sL[r * ldl + q + 1] * sR[(q + 1) * ldr + t]; }
          if (q < r) acc0 -= sL[r * ldl + q] * sR[q * ldr + t];
          sR[r * ldr + t] = (acc0 + acc1) * sInv[r];
        }
      }
      __syncthreads();
      for (int e = tid; e < nb * nb; e += kBcrThreads) {
        const int u = e / nb, v = e - u * nb;
        N[pl.oL + e] = (v <= u) ? sL[u * ldl + v] : 0.0;
        N[pl.oYa + e] = sR[u * ldr + v];
        N[pl.oYc + e] = sR[u * ldr + nb + v];
      }
      for (int e = tid; e < nb * m; e += kBcrThreads) {
        const int u = e / m, r = e - u * m;
        N[pl.oYf + e] = sR[u * ldr + 2 * nb + r];                      // Y_f [nb][m]
      }
      for (int u = tid; u < nb; u += kBcrThreads) N[pl.oy + u] = sR[u * ldr + 2 * nb + m];
      __syncthreads();
    }
    bcr_grid_sync(barrier, &s_phase);
    // phase 2: even nodes e = 0, 2s, 4s, ...: Schur complements of the eliminated neighbours e - s and e + s
    for (int e0 = 2 * s * cta; e0 < nsb; e0 += 2 * s * nctas) {
      double* N = node(e0);
      const int i1 = e0 - s, i2 = e0 + s;
      const bool h1 = i1 >= 0, h2 = i2 < nsb;
      if (!h1 && !h2) continue;
      const bool hnext = e0 + 2 * s < nsb;
      // stage: P1 = Y_c(i1) [nb][nb], P2 = Y_a(i2), P3 = Y_c(i2), Q1 = Y_f(i1) [nb][m], Q2 = Y_f(i2), y1, y2
      double* P1 = s_bcr; double* P2 = P1 + nb * nb; double* P3 = P2 + nb * nb;
      double* Q1 = P3 + nb * nb; double* Q2 = Q1 + nb * m; double* y1 = Q2 + nb * m; double* y2 = y1 + nb;
      for (int e = tid; e < nb * nb; e += kBcrThreads) {
        P1[e] = h1 ? node(i1)[pl.oYc + e] : 0.0;
        P2[e] = h2 ? node(i2)[pl.oYa + e] : 0.0;
        P3[e] = (h2 && hnext) ? node(i2)[pl.oYc + e] : 0.0;
      }
      for (int e = tid; e < nb * m; e += kBcrThreads) { Q1[e] = h1 ? node(i1)[pl.oYf + e] : 0.0; Q2[e] = h2 ? node(i2)[pl.oYf + e] : 0.0; }
      for (int u = tid; u < nb; u += kBcrThreads) { y1[u] = h1 ? node(i1)[pl.oy + u] : 0.0; y2[u] = h2 ? node(i2)[pl.oy + u] : 0.0; }
      __syncthreads();
      for (int e = tid; e < nb * nb; e += kBcrThreads) {
        const int u = e / nb, v = e - u * nb;
        double d = 0.0, bnew = 0.0;
        for (int q = 0; q < nb; ++q) {
          d += P1[q * nb + u] * P1[q * nb + v] + P2[q * nb + u] * P2[q * nb + v];
          bnew += P3[q * nb + u] * P2[q * nb + v];   // (Y_c^T Y_a)[u][v]: rows = dofs of e + 2s, columns = dofs of e
        }
        N[pl.oD + e] -= d;
        N[pl.oB + e] = -bnew;
      }
      for (int e = tid; e < m * nb; e += kBcrThreads) {
        const int r = e / nb, v = e - r * nb;
        double f = 0.0;
        for (int q = 0; q < nb; ++q) f += Q1[q * m + r] * P1[q * nb + v] + Q2[q * m + r] * P2[q * nb + v];
        N[pl.oF + e] -= f;
      }
      for (int v = tid; v < nb; v += kBcrThreads) {
        double g = 0.0;
        for (int q = 0; q < nb; ++q) g += P1[q * nb + v] * y1[q] + P2[q * nb + v] * y2[q];
        N[pl.ob + v] -= g;
      }
      __syncthreads();
    }
    bcr_grid_sync(barrier, &s_phase);
  }
  // ---- top node 0: factor, then the corner ----
  if (cta == 0) {
    double* N = node(0);
    for (int e = tid; e < nb * nb; e += kBcrThreads) { const int u = e / nb, v = e - u * nb; sL[u * ldl + v] = N[pl.oD + e]; }
    for (int e = tid; e < m * nb; e += kBcrThreads) { const int r = e / nb, u = e - r * nb; sR[u * ldr + r] = N[pl.oF + e]; }
    for (int u = tid; u < nb; u += kBcrThreads) sR[u * ldr + m] = N[pl.ob + u];
    __syncthreads();
    bcr_chol(sL, ldl, nb, sInv, &s_ok);
    for (int t = tid; t < m + 1; t += kBcrThreads)
      for (int r = 0; r < nb; ++r) {
        double acc = sR[r * ldr + t];
        for (int q = 0; q < r; ++q) acc -= sL[r * ldl + q] * sR[q * ldr + t];
        sR[r * ldr + t] = acc * sInv[r];
      }
    __syncthreads();
    for (int e = tid; e < nb * nb; e += kBcrThreads) { const int u = e / nb, v = e - u * nb; N[pl.oL + e] = (v <= u) ? sL[u * ldl + v] : 0.0; }
    for (int e = tid; e < nb * m; e += kBcrThreads) { const int u = e / m, r = e - u * m; N[pl.oYf + e] = sR[u * ldr + r]; }
    for (int u = tid; u < nb; u += kBcrThreads) N[pl.oy + u] = sR[u * ldr + m];
  }
  bcr_grid_sync(barrier, &s_phase);
  // corner update partials: CTA c sums Y_f^T [Y_f | y] over its nodes (fixed assignment, fixed order)
  {
    double* part = ws + pl.oCCp + static_cast<long long>(cta) * (m + 1) * m;
    for (int e = tid; e < (m + 1) * m; e += kBcrThreads) {
      const int r = e / m, q = e - r * m;   // r == m: rhs row
      double acc = 0.0;
      if (r == m || q <= r) {
        for (int j = cta; j < nsb; j += nctas) {
          const double* Yf = node(j) + pl.oYf;
          const double* yv = node(j) + pl.oy;
          for (int u = 0; u < nb; ++u) acc += Yf[u * m + q] * (r < m ? Yf[u * m + r] : yv[u]);
        }
      }
      part[e] = acc;
    }
  }
  bcr_grid_sync(barrier, &s_phase);
  double* xa = ws + pl.oXa;
  if (cta == 0) {
    // corner: C' = C - sum of partials (CTA order), Cholesky (m x m), forward + backward substitution
    double* CCs = s_bcr;
    const int ldc = m | 1;
    double* cinv = CCs + (m + 1) * ldc;
    double* rhs = cinv + m + (m & 1);
    const double* CC = ws + pl.oCC;
    for (int e = tid; e < (m + 1) * m; e += kBcrThreads) {
      const int r = e / m, q = e - r * m;
      double v = CC[e];
      for (int c = 0; c < nctas; ++c) v -= ws[pl.oCCp + static_cast<long long>(c) * (m + 1) * m + e];
      CCs[r * ldc + q] = v;
    }
    __syncthreads();
    bcr_chol(CCs, ldc, m, cinv, &s_ok);
    if (tid < 32) {
      // y = L^-1 rhs (row m), x = L^-T y: one warp, sequential in q (m is a few dozen)
      if (tid == 0) {
        for (int r = 0; r < m; ++r) {
          double acc = CCs[m * ldc + r];
          for (int q = 0; q < r; ++q) acc -= CCs[r * ldc + q] * rhs[q];
          rhs[r] = acc * cinv[r];
        }
        for (int r = m - 1; r >= 0; --r) {
          double acc = rhs[r];
          for (int q = r + 1; q < m; ++q) acc -= CCs[q * ldc + r] * rhs[q];
          rhs[r] = acc * cinv[r];
        }
      }
      __syncwarp();
      for (int r = tid; r < m; r += 32) { xa[r] = rhs[r]; x_out[np + r] = rhs[r]; }
    }
  }
  bcr_grid_sync(barrier, &s_phase);
  // ---- way back: x_i = L^-T (y - Y_a x_a - Y_c x_c - Y_f x_arrow); top node first, then level by level ----
  auto backsolve_node = [&](int i, int a, int c) {
    const double* N = node(i);
    double* rv = s_bcr;                 // [nb]
    double* sLt = s_bcr + nb + (nb & 1); // L [nb][ldl]
    for (int e = tid; e < nb * nb; e += kBcrThreads) { const int u = e / nb, v = e - u * nb; sLt[u * ldl + v] = N[pl.oL + e]; }
    for (int u = tid; u < nb; u += kBcrThreads) {
      double acc = N[pl.oy + u];
      for (int r = 0; r < m; ++r) acc -= N[pl.oYf + u * m + r] * xa[r];
      if (a >= 0) for (int v = 0; v < nb; ++v) { const int col = nb * a + v; if (col < np) acc -= N[pl.oYa + u * nb + v] * x_out[col]; }
      if (c >= 0) for (int v = 0; v < nb; ++v) { const int col = nb * c + v; if (col < np) acc -= N[pl.oYc + u * nb + v] * x_out[col]; }
      rv[u] = acc;
    }
    __syncthreads();
    if (tid == 0) {
      for (int r = nb - 1; r >= 0; --r) {
        double acc = rv[r];
        for (int q = r + 1; q < nb; ++q) acc -= sLt[q * ldl + r] * rv[q];
        rv[r] = acc / sLt[r * ldl + r];
      }
    }
    __syncthreads();
    for (int u = tid; u < nb; u += kBcrThreads) { const int col = nb * i + u; if (col < np) x_out[col] = rv[u]; }
    __syncthreads();
  };
  if (cta == 0) backsolve_node(0, -1, -1);
  bcr_grid_sync(barrier, &s_phase);
  for (int lv = pl.levels - 1; lv >= 0; --lv) {
    const int s = 1 << lv;
    for (int i = s + 2 * s * cta; i < nsb; i += 2 * s * nctas) backsolve_node(i, i - s, (i + s < nsb) ? i + s : -1);
    if (lv > 0) bcr_grid_sync(barrier, &s_phase);
  }
  if (tid == 0 && !s_ok) atomicExch(spd_flag, 0);
}

}  // namespace hb
