// One-sided (Hestenes) Jacobi SVD on tiny matrices held in registers, fp64.
//
// Why not eig(A^T A): the DLT system has entries ~4e5 next to ~1 (P = K[R|t] in pixels x
// millimetres), so forming A^T A squares a condition number of ~1e6.
// Rotating the columns of A directly keeps the relative accuracy of the small singular
// vectors at ~eps*cond(A) in fp64, far below the 1e-2 mm parity gate.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#if defined(CDR_JACOBI_STATS)
static long long cdr_jacobi_sweeps = 0;   // host-side instrumentation (tests/host)
#endif

namespace cdr {

// 1/x and 1/sqrt(x) for the rotation parameters.  Device: the hardware's 20-bit fp64 seed (MUFU.RCP64H / RSQ64H, full
// exponent range) + two Newton steps — ~4x shorter dependent chains than DDIV / DSQRT, which dominated the latency of
// the per-joint solve (one thread per joint: the chain IS the kernel time).  The results are good to a few ulp; the
// Jacobi iteration is self-correcting (an angle off by 1e-15 leaves an off-diagonal the next sweep removes, and a
// rotation scaled by 1 + 1e-16 scales both columns of G and V alike — the de-homogenised point only sees ratios).
__host__ __device__ __forceinline__ double jacobi_rcp(double x) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
#else
  return 1.0 / x;
#endif
}
__host__ __device__ __forceinline__ double jacobi_rsqrt(double x) {
#ifdef __CUDA_ARCH__
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  y = fma(y, fma(-hx * y, y, 0.5), y);
  y = fma(y, fma(-hx * y, y, 0.5), y);
  return y;
#else
  return 1.0 / sqrt(x);
#endif
}

// Orthogonalise the C columns of G (R x C, column access G[r][c]) in place by plane
// rotations from the right, accumulating them into V (C x C, starts as identity).
// On exit G = U*Sigma (column norms are the singular values) and A_original * V = G.
template <int R, int C>
__host__ __device__ __forceinline__ void jacobi_onesided(double (&G)[R][C], double (&V)[C][C]) {
#pragma unroll
  for (int i = 0; i < C; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;

  const double tol = 1e-15;
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool rotated = false;
#pragma unroll
    for (int p = 0; p < C - 1; ++p) {
#pragma unroll
      for (int q = p + 1; q < C; ++q) {
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          alpha = fma(G[r][p], G[r][p], alpha);
          beta = fma(G[r][q], G[r][q], beta);
          gamma = fma(G[r][p], G[r][q], gamma);
        }
        // skip when the pair is already orthogonal to working precision (also covers
        // zero columns: gamma == 0)
        if (gamma * gamma > (tol * tol) * (alpha * beta) && gamma != 0.0) {
          rotated = true;
          // tan(theta) = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = (beta - alpha) / (2 gamma), in the
          // division-free form t = sign(d) * 2 gamma / (|d| + sqrt(d^2 + 4 gamma^2)): one rsqrt + one reciprocal
          const double d = beta - alpha, g2 = 2.0 * gamma;
          const double h2 = fma(d, d, g2 * g2);
          const double h = h2 * jacobi_rsqrt(h2);
          const double t = (d < 0.0 ? -g2 : g2) * jacobi_rcp(fabs(d) + h);
          const double c = jacobi_rsqrt(fma(t, t, 1.0));
          const double s = c * t;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const double gp = G[r][p], gq = G[r][q];
            G[r][p] = c * gp - s * gq;
            G[r][q] = s * gp + c * gq;
          }
#pragma unroll
          for (int r = 0; r < C; ++r) {
            const double vp = V[r][p], vq = V[r][q];
            V[r][p] = c * vp - s * vq;
            V[r][q] = s * vp + c * vq;
          }
        }
      }
    }
    if (!rotated) break;
#if defined(CDR_JACOBI_STATS) && !defined(__CUDA_ARCH__)
    ++cdr_jacobi_sweeps;
#endif
  }
}

// Null-space direction of a 4x4 DLT system: right singular vector of the smallest
// singular value, de-homogenised (x/w, y/w, z/w).  No guard on w ~ 0: like the reference
// (models/cdrnet.py:175-177) this returns inf/nan for points at infinity.
__host__ __device__ __forceinline__ void dlt_solve4(double (&A)[4][4], double& x, double& y, double& z) {
  double V[4][4];
  jacobi_onesided<4, 4>(A, V);
  double best = INFINITY;
  int k = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    double n2 = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) n2 = fma(A[r][c], A[r][c], n2);
    if (n2 < best) { best = n2; k = c; }
  }
  double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {   // select without dynamic register indexing
    if (c == k) { v0 = V[0][c]; v1 = V[1][c]; v2 = V[2][c]; v3 = V[3][c]; }
  }
  x = v0 / v3;
  y = v1 / v3;
  z = v2 / v3;
}

// Rows of the reference's DLT system for one view (models/cdrnet.py:169-171):
//   [u*P[2] - P[0] ; v*P[2] - P[1]]   with P row-major 3x4.
template <typename TP>
__host__ __device__ __forceinline__ void dlt_rows(const TP* __restrict__ P, double u, double v,
                                         double (&A)[4][4], int row0) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const double p2 = (double)P[8 + c];
    A[row0][c] = u * p2 - (double)P[c];
    A[row0 + 1][c] = v * p2 - (double)P[4 + c];
  }
}

// Backward of the two-view DLT (models/cdrnet.py:151-179) for one joint: given dL/dX for X = v[:3] / v[3], v the
// right singular vector of the smallest singular value of A(u_l, v_l, u_r, v_r), returns dL/d(u_l, v_l, u_r, v_r).
// With M = A^T A: dv = -(M - s^2 I)^+ dM v (the gauge of torch.svd's backward; X does not see the sign of v), so
// for w = -(M - s^2 I)^+ g_v:  dL/dA = (A v) w^T + (A w) v^T, and dA[row]/d(coordinate) is the row P_view[2].
// The one-sided Jacobi yields G = A V (columns sigma_c u_c) and V at once.
template <typename TP>
__host__ __device__ __forceinline__ void dlt_backward4(const TP* __restrict__ Pl, const TP* __restrict__ Pr, double ul,
                                                       double vl, double ur, double vr, const double (&gX)[3],
                                                       double (&g_kp)[4]) {
  double G[4][4], V[4][4];
  dlt_rows(Pl, ul, vl, G, 0);
  dlt_rows(Pr, ur, vr, G, 2);
  jacobi_onesided<4, 4>(G, V);
  double s2[4], best = INFINITY;
  int k = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    s2[c] = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) s2[c] = fma(G[r][c], G[r][c], s2[c]);
    if (s2[c] < best) { best = s2[c]; k = c; }
  }
  double v[4] = {0, 0, 0, 0}, Av[4] = {0, 0, 0, 0};
#pragma unroll
  for (int c = 0; c < 4; ++c)
    if (c == k) {
#pragma unroll
      for (int r = 0; r < 4; ++r) { v[r] = V[r][c]; Av[r] = G[r][c]; }
    }
  const double iw = 1.0 / v[3];
  const double gv[4] = {gX[0] * iw, gX[1] * iw, gX[2] * iw, -(gX[0] * v[0] + gX[1] * v[1] + gX[2] * v[2]) * iw * iw};
  double w[4] = {0, 0, 0, 0}, Aw[4] = {0, 0, 0, 0};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c == k) continue;
    double dot = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) dot = fma(V[r][c], gv[r], dot);
    const double coef = -dot / (s2[c] - best);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      w[r] = fma(coef, V[r][c], w[r]);
      Aw[r] = fma(coef, G[r][c], Aw[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const TP* P2 = (r < 2 ? Pl : Pr) + 8;
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < 4; ++c) acc = fma(Av[r] * w[c] + Aw[r] * v[c], (double)P2[c], acc);
    g_kp[r] = acc;
  }
}

// Moore-Penrose pseudo-inverse of a row-major 3x4 matrix with torch.linalg.pinv's
// semantics (models/cdrnet.py:236-237): singular values <= rtol*sigma_max are dropped.
// Works on G = P^T (4x3): G V = U Sigma, so pinv(P) = sum_i G_i V_i^T / sigma_i^2 (4x3).
template <typename TIn, typename TOut>
__host__ __device__ __forceinline__ void pinv_3x4(const TIn* __restrict__ P, double rtol,
                                                  TOut* __restrict__ out) {
  double G[4][3], V[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) G[c][r] = (double)P[r * 4 + c];
  jacobi_onesided<4, 3>(G, V);
  double s2[3], smax2 = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    s2[i] = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) s2[i] = fma(G[r][i], G[r][i], s2[i]);
    smax2 = s2[i] > smax2 ? s2[i] : smax2;
  }
  const double cut2 = rtol * rtol * smax2;
  double inv[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) inv[i] = (s2[i] > cut2 && s2[i] > 0.0) ? 1.0 / s2[i] : 0.0;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      double acc = 0.0;
#pragma unroll
      for (int i = 0; i < 3; ++i) acc = fma(G[a][i] * inv[i], V[b][i], acc);
      out[a * 3 + b] = (TOut)acc;
    }
}

}  // namespace cdr
