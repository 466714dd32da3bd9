// zsolve_seg.cuh — warp-parallel tridiagonal z solve in ONE pass over HBM.
//
// Replaces the two sweeps of solmpj (src/modpois.f90:1120-1162), which cross memory twice, and is the north_star's
// "warp-parallel reduction in z".  A CTA owns TW neighbouring columns of the (i,j) plane (one TW*8-byte row per level)
// and NSEG = K / L segments; the TW threads of segment s hold levels s*L .. s*L+L-1 of the TW columns in registers
// (thread = column, so every global access is a full coalesced row and nothing is transposed).  With the tabulated
// factors z_k (k_zfactor) both sweeps are affine recurrences,
//     forward   y_k = alpha_k y_{k-1} + x_k z_k,   alpha_k = -a_k z_k
//     backward  x_k = y_k + delta_k x_{k+1},       delta_k = -c_k z_k      (c_{K-1} = 0),
// so a segment acts on its successor through a pair (product of its alpha's, its local result with zero inflow).
// Every segment runs its L levels with zero inflow, the NSEG pairs meet in shared memory, every segment folds the
// pairs of its predecessors (<= NSEG-1 FMAs) into its true inflow and adds inflow * running product to its levels:
// the dependent chain is 4 L + 2 NSEG operations instead of 2 K; x is read once and written once.  Same factors as
// the streaming kernel, other association of the sums than solmpj: differences are rounding (tests: 1e-10 relative).
//
// Measured at 256^3 (profiles/r2_ab7_zseg.jsonl, r2_zseg_ncu.txt): 16 levels per thread, 16-column tiles, two CTAs of
// 256 threads per SM: 62.9 us under ncu against 93.8 us for the streaming kernel, DRAM traffic 247 MB against 404 MB;
// poisson_core 0.367 -> 0.349 ms inside the substep.  Wider tiles (32 columns: 0.362), shorter segments (8 levels:
// 0.369), three or four CTAs per SM with the factors parked in shared memory (0.365 / 0.375) were all slower.
#pragma once
#include "common.cuh"

namespace udg {

template <int L, int TW, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
k_zsolve_seg(Geo g, int nxh, int nyh, int nseg, double *__restrict__ x, const double *__restrict__ zt,
             const double *__restrict__ a, const double *__restrict__ c) {
  extern __shared__ double zs_sm[];                       // [4][nseg][TW] pairs, then -a[K], -c[K]
  const int col = threadIdx.x % TW, s = threadIdx.x / TW;
  const int K = g.ktot;
  const long long plane = (long long)g.imax * g.jmax, tk = (long long)nxh * nyh;
  const long long q0 = (long long)blockIdx.x * TW + col;
  const bool act = q0 < plane;
  const long long q = act ? q0 : plane - 1;
  const int i = (int)(q % g.imax), j = (int)(q / g.imax);
  const int ig = g.i0g + i;
  const int ix = g.xalt ? (ig == 0 ? 0 : ig == 1 ? nxh - 1 : ig >> 1) : (ig + 1) >> 1;   // x slot -> distinct eigenvalue
  const int jy = (g.j0g + j + 1) >> 1;                                                   // packed y slot -> distinct eigenvalue
  const int k0 = s * L;
  double *xp = x + q + (long long)k0 * plane;
  const double *zp = zt + (long long)jy * nxh + ix + (long long)k0 * tk;
  double xr[L], zr[L];
#pragma unroll
  for (int u = 0; u < L; u++) { xr[u] = xp[u * plane]; zr[u] = __ldg(zp + u * tk); }
  double *sA = zs_sm, *sB = sA + nseg * TW, *sC = sB + nseg * TW, *sD = sC + nseg * TW;
  double *sa = sD + nseg * TW + k0, *sc = sa + K;      // this segment's -a_k, -c_k (shared: no registers, no hoisting)
  for (int k = threadIdx.x; k < K; k += blockDim.x) { sD[nseg * TW + k] = -__ldg(a + k); sD[nseg * TW + K + k] = -__ldg(c + k); }
  __syncthreads();
  // forward, zero inflow
  double y = 0., P = 1.;
#pragma unroll
  for (int u = 0; u < L; u++) {
    const double al = sa[u] * zr[u];
    y = fma(al, y, xr[u] * zr[u]);
    xr[u] = y;
    P *= al;
  }
  sA[s * TW + col] = P;
  sB[s * TW + col] = y;
  __syncthreads();
  double yin = 0.;
  for (int t = 0; t < s; t++) yin = fma(sA[t * TW + col], yin, sB[t * TW + col]);
  P = 1.;
#pragma unroll
  for (int u = 0; u < L; u++) {
    P *= sa[u] * zr[u];
    xr[u] = fma(P, yin, xr[u]);
  }
  // backward, zero inflow from above
  double xh = 0., Q = 1.;
#pragma unroll
  for (int u = L - 1; u >= 0; u--) {
    const double de = sc[u] * zr[u];
    xh = fma(de, xh, xr[u]);
    xr[u] = xh;
    Q *= de;
  }
  sC[s * TW + col] = Q;
  sD[s * TW + col] = xh;
  __syncthreads();
  double xin = 0.;
  for (int t = nseg - 1; t > s; t--) xin = fma(sC[t * TW + col], xin, sD[t * TW + col]);
  Q = 1.;
#pragma unroll
  for (int u = L - 1; u >= 0; u--) {
    Q *= sc[u] * zr[u];
    xr[u] = fma(Q, xin, xr[u]);
  }
  if (act) {
#pragma unroll
    for (int u = 0; u < L; u++) xp[u * plane] = xr[u];
  }
}

}  // namespace udg
