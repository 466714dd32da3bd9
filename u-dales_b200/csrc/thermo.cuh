// thermo.cuh — temperature on the resident path, dry: ltempeq with iadv_thl = cd2, lbuoyancy.
//   advecc_2nd + diffc on thl0            src/modadvection.f90:67-69, src/modsubgrid.f90:146  -> k_scalar_tend<2,..> on a
//                                         momentum-halo geometry (scalar_v1.cuh), no kernel of its own
//   bottom, fixed temperature flux        src/modibm.f90:2033-2046                            -> k_bottom_scalar (flux argument)
//   forces: buoyancy, radiative tendency  src/modforces.f90:70-83, 103-109                    -> k_buoyancy, k_tend_add_profile
//   boundary / closurebc: top of thl      src/modboundary.f90:208-221, 417-420, 1494-1517     -> k_thl_top
//   thermodynamics                        src/modthermodynamics.f90:55-121 (diagfld :270, calc_halflev :508-526, calthv dry
//                                         branch :213-235, thvh :76-90)                        -> k_thermo_partial, k_thermo_final
//   ibmnorm / ibmwallfun on temperature   src/modibm.f90:714-722, 936-987, 1225               -> k_ibm_solid_scalar, k_ibm_diffcorr_c
//                                                                                                (ibm.cuh) + k_ibm_advecc2nd_corr
// thl0h / thv0h are never stored: in the dry case thv0h = thl0h is the dzf-weighted mean of two levels of thl0
// (:521), which the buoyancy kernel and the slab means evaluate where they need it; thl0 does not change between
// thermodynamics() and the next forces() (src/program.f90:158-212), so the values are the reference's.
#pragma once
#include "common.cuh"

namespace udg {

// thl0h(i,j,k), k >= kb+1 (src/modthermodynamics.f90:521); c = offset of (i,j,k) in the momentum-halo array
__device__ __forceinline__ double thl_half(const Geo &g, const double *__restrict__ thl0, long long c, int k) {
  return (thl0[c] * g.dzf[k - 1] + thl0[c - g.pk] * g.dzf[k]) / (2 * g.dzh[k]);
}

// wp(i,j,k) += grav (thv0h(i,j,k) - thvh(k)) / thvh(k), k = kb+1 .. ke   (src/modforces.f90:73-81; blockIdx.z = k - 2)
__global__ void __launch_bounds__(256) k_buoyancy(Geo g, double grav, const double *__restrict__ thl0, const double *__restrict__ thvh /* index k */,
                                                  double *__restrict__ wp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 2;
  if (i > g.imax || j > g.jmax) return;
  const long long t = offT(g, i, j, k);
  const double th = thvh[k];
  wp[t] = wp[t] + grav * (thl_half(g, thl0, offF(g, i, j, k), k) - th) / th;
}

// thlp(i,j,k) += thlpcar(k), k = kb .. ke   (src/modforces.f90:103-109)
__global__ void __launch_bounds__(256) k_tend_add_profile(Geo g, const double *__restrict__ prof /* index k */, double *__restrict__ tp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  if (i > g.imax || j > g.jmax) return;
  const long long t = offT(g, i, j, k);
  tp[t] = tp[t] + prof[k];
}

// top ghost level of thlm and thl0 over the whole halo'd plane: fluxtop(field, ekh, wttop) (mode 1; zero flux = copy)
// or valuetop(field, thl_top) (mode 2)   (src/modboundary.f90:1494-1517)
__global__ void k_thl_top(Geo g, int mode, double flux, double val, const double *__restrict__ ekh, double *__restrict__ thl0,
                          double *__restrict__ thlm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1 - g.ih;
  const int j = blockIdx.y + 1 - g.jh;
  if (i > g.imax + g.ih) return;
  const int K = g.ktot;
  const long long c = offF(g, i, j, K), cp = c + g.pk;
  if (mode == 2) {
    thlm[cp] = 2 * val - thlm[c];
    thl0[cp] = 2 * val - thl0[c];
  } else if (fabs(flux) <= 1.e-10) {
    thlm[cp] = thlm[c];
    thl0[cp] = thl0[c];
  } else {
    const double add = g.dzh[K + 1] * flux / (g.dzhi[K + 1] * (0.5 * (g.dzf[K] * ekh[cp] + g.dzf[K + 1] * ekh[c])));
    thlm[cp] = thlm[c] + add;
    thl0[cp] = thl0[c] + add;
  }
}

// plane sums for thermodynamics(), levels k = 1 .. K+1, in a fixed order (block b of level k sums its strided share):
//   slot 0: sum thl0 * mask_c     (diagfld: thl0av, avexy_ibm with IIc)
//   slot 1: sum thl0              (the kb fallback of avexy_ibm when every cell of kb is solid, src/modmpi.f90:649-652)
//   slot 2: sum thl0h * mask_w    (thvh = avexy_ibm(thv0h, IIw); thl0h(kb) = thls)
constexpr int TH_NBLK = 64;
__global__ void __launch_bounds__(256) k_thermo_partial(Geo g, double thls, const double *__restrict__ thl0, const double *__restrict__ mask_c,
                                                        const double *__restrict__ mask_w, double *__restrict__ part /* [3][K+1][TH_NBLK] */) {
  const int k = blockIdx.y + 1, b = blockIdx.x, K1 = g.ktot + 1;
  double s0 = 0., s1 = 0., s2 = 0.;
  // block b takes rows j = b+1, b+1+TH_NBLK, ..; threads run along i (coalesced); the order of the additions is fixed
  for (int j = b + 1; j <= g.jmax; j += TH_NBLK) {
    const long long row = offF(g, 1, j, k);
    for (int i = threadIdx.x; i < g.imax; i += blockDim.x) {
      const long long c = row + i;
      const double v = thl0[c];
      s0 += v * (mask_c ? mask_c[c] : 1.0);
      s1 += v;
      const double vh = k == 1 ? thls : thl_half(g, thl0, c, k);
      s2 += vh * (mask_w ? mask_w[c] : 1.0);
    }
  }
  __shared__ double sh[3][256];
  sh[0][threadIdx.x] = s0; sh[1][threadIdx.x] = s1; sh[2][threadIdx.x] = s2;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if ((int)threadIdx.x < o)
      for (int s = 0; s < 3; s++) sh[s][threadIdx.x] += sh[s][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x < 3) part[((long long)threadIdx.x * K1 + (k - 1)) * TH_NBLK + b] = sh[threadIdx.x][0];
}
__global__ void k_thermo_reduce(int n, const double *__restrict__ part, double *__restrict__ sums) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  double s = 0.;
  for (int b = 0; b < TH_NBLK; b++) s += part[(long long)q * TH_NBLK + b];
  sums[q] = s;
}
// profiles from the (cross-rank) sums: thl0av(k), thvh(k), k = 1 .. K+1 (tables indexed by k), and the value ibmnorm
// gives solid temperature points, sum(thl0av(kb:ke) dzf(kb:ke)) / zh(ke+1) (src/modibm.f90:715).
// cnt_c / cnt_w: fluid points per level of mask_c / mask_w over all ranks (IIcs, IIws).  One block of 256 threads: a
// thread per level (a single thread walking the levels pays one dependent global load per level, ~100 us at K = 256),
// the weighted sum through a fixed shared-memory tree.
__global__ void __launch_bounds__(256) k_thermo_final(int K, const double *__restrict__ sums /* [3][K+1] */, const double *__restrict__ cnt_c,
                                                      const double *__restrict__ cnt_w, const double *__restrict__ dzf, double zhtop,
                                                      double *__restrict__ thl0av, double *__restrict__ thvh, double *__restrict__ solid_val) {
  const int K1 = K + 1;
  const double eps1 = 1.e-10;
  __shared__ double sh[256];
  double ws = 0.;
  for (int k = threadIdx.x + 1; k <= K1; k += blockDim.x) {
    double d = cnt_c[k - 1], a = sums[k - 1];
    if (k == 1 && d == 0.) { a = sums[K1 + k - 1]; d = cnt_c[K - 1]; }
    const double av = d == 0. ? -999. : a / d;
    thl0av[k] = av;
    const double dw = cnt_w[k - 1];
    thvh[k] = dw == 0. ? -999. : sums[2 * K1 + k - 1] / dw;
    if (k <= K) ws += av * dzf[k];
  }
  sh[threadIdx.x] = ws;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    thvh[1] = thl0av[1];                                   // :87  th0av(kb) (1 + 0 - 0), dry
    if (K1 >= 2 && fabs(thvh[2]) < eps1) thvh[2] = thl0av[2];   // :88-90
    *solid_val = sh[0] / zhtop;
  }
}

// Buoyancy correction of the Vreman eddy viscosity for stable stratification (lbuoycorr, src/modsubgrid.f90:332-354), as a
// pass over the interior after the closure kernel: ekm there is nu_t + numol and ekh = nu_t prandtli + numol prandtlmoli
// (:356-360), so nu_t = ekm - numol is scaled by sqrt(1 - min(max(Rig, 0), Rigc) / Rigc) and both are rebuilt.  dthvdz is
// calthv's (dry: centred difference of thl0, 0 at kb, floored at +-eps1; src/modthermodynamics.f90:213-235) evaluated in
// place: thl0 has not changed since the last thermodynamics().
__global__ void __launch_bounds__(256) k_vreman_buoycorr(Geo g, double grav, double Rigc, const double *__restrict__ u0, const double *__restrict__ v0,
                                                         const double *__restrict__ thl0, double *__restrict__ ekm, double *__restrict__ ekh) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  if (i > g.imax || j > g.jmax) return;
  const long long c = offF(g, i, j, k), sj = g.pi, sk = g.pk;
  const double eps1 = 1.e-10;
  double dth = k >= 2 ? (thl0[c + sk] - thl0[c - sk]) / (g.dzh[k + 1] + g.dzh[k]) : 0.;
  if (fabs(dth) < eps1) dth = copysign(eps1, dth);
  const double dzs = g.dzh[k + 1] + g.dzh[k];
  const double du0dz = 0.5 * ((u0[c + sk] + u0[c + 1 + sk]) - (u0[c - sk] + u0[c + 1 - sk])) / dzs;
  const double dv0dz = 0.5 * ((v0[c + sk] + v0[c + sj + sk]) - (v0[c - sk] + v0[c + sj - sk])) / dzs;
  const double Rig = ((grav / thl0[c]) * dth) / (du0dz * du0dz + dv0dz * dv0dz + 1.e-10);
  const double e = (ekm[c] - g.numol) * sqrt(1.0 - fmin(fmax(Rig, 0.0), Rigc) / Rigc);
  ekh[c] = e * g.prandtli + g.numol * g.prandtlmoli;
  ekm[c] = e + g.numol;
}

// advecc2nd_corr_liberal (src/modibm.f90:936-987) on a momentum-halo scalar: one thread per fluid-boundary point of c
__global__ void k_ibm_advecc2nd_corr(Geo g, int n, const int *__restrict__ pts, const double *__restrict__ mk, const double *__restrict__ u0,
                                     const double *__restrict__ v0, const double *__restrict__ w0, const double *__restrict__ var,
                                     double *__restrict__ rhs) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double eps1 = 1.e-10;
  const int i = pts[3 * p], j = pts[3 * p + 1], k = pts[3 * p + 2];
#define V(a, b, c) var[offF(g, a, b, c)]
#define MK(a, b, c) (fabs(mk[offF(g, a, b, c)]) < eps1)
  double t = rhs[offT(g, i, j, k)];
  const double vc = V(i, j, k);
  if (MK(i + 1, j, k)) { const double u = u0[offF(g, i + 1, j, k)]; t = t + u * (V(i + 1, j, k) + vc) * g.dxi5 - u * (vc + vc) * g.dxi5; }
  if (MK(i - 1, j, k)) { const double u = u0[offF(g, i, j, k)];     t = t - u * (V(i - 1, j, k) + vc) * g.dxi5 + u * (vc + vc) * g.dxi5; }
  if (MK(i, j + 1, k)) { const double v = v0[offF(g, i, j + 1, k)]; t = t + v * (V(i, j + 1, k) + vc) * g.dyi5 - v * (vc + vc) * g.dyi5; }
  if (MK(i, j - 1, k)) { const double v = v0[offF(g, i, j, k)];     t = t - v * (V(i, j - 1, k) + vc) * g.dyi5 + v * (vc + vc) * g.dyi5; }
  if (MK(i, j, k + 1)) {
    const double w = w0[offF(g, i, j, k + 1)];
    t = t + w * (V(i, j, k + 1) * g.dzf[k] + vc * g.dzf[k + 1]) * g.dzhi[k + 1] * g.dzfi5[k]
          - w * (vc * g.dzf[k] + vc * g.dzf[k + 1]) * g.dzhi[k + 1] * g.dzfi5[k];
  }
  if (MK(i, j, k - 1)) {
    const double w = w0[offF(g, i, j, k)];
    t = t - w * (V(i, j, k - 1) * g.dzf[k] + vc * g.dzf[k - 1]) * g.dzhi[k] * g.dzfi5[k]
          + w * (vc * g.dzf[k] + vc * g.dzf[k - 1]) * g.dzhi[k] * g.dzfi5[k];
  }
  rhs[offT(g, i, j, k)] = t;
#undef V
#undef MK
}

}  // namespace udg
