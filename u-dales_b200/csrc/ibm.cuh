// ibm.cuh — immersed-boundary masking on the device (next tier): the sparse point-list routines of
// src/modibm.f90 that sit on the substep path between subgrid and poisson:
//   solid        :748-826   velocities / tendencies at solid points (and scalars: average of the fluid neighbours)
//   ibmnorm      :697-745   solid() for um/up, vm/vp, wm/wp and every scalar
//   diffu_corr, diffv_corr, diffw_corr, diffc_corr  :990-1164   cancel the subgrid flux through solid neighbours
//   masks        :153-192   real masks (1 fluid, 0 solid) incl. ground level, halo-exchanged
// Point lists are local 1-based (i,j,k) triples (solid_info%solpts_loc / bound_info%bndpts_loc).  One thread per
// point: every routine only writes its own point and reads fluid neighbours, so list order does not matter.
// Operand order is the reference's: results differ from the oracle's only by FMA contraction.
#pragma once
#include "common.cuh"

namespace udg {

__global__ void k_ibm_mask_init(Geo g, double *__restrict__ mk, int is_w) {
  const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long n = g.pk * (g.ktot + 2 * g.kh);
  if (q >= n) return;
  const int lev = (int)(q / g.pk);   // storage level 0 = Fortran kb-kh
  mk[q] = (lev == 0 || (is_w && lev == 1)) ? 0. : 1.;
}
__global__ void k_ibm_mask_solid(Geo g, int n, const int *__restrict__ pts, double *__restrict__ mk) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  mk[offF(g, pts[3 * p], pts[3 * p + 1], pts[3 * p + 2])] = 0.;
}

// solid() without mask (:762-770).  pend != nullptr: a per-level table (forces / masscorr, index k) is still to be
// subtracted from this tendency inside the fused tderive+integrate kernel; the solid point gets the table value so that
// the subtraction leaves exactly the 0 the reference stores (x - x), and every other reader of the tendency in between
// (fillps: differences within a level) sees all points of a level carrying the same offset.
__global__ void k_ibm_solid_mom(Geo g, int n, const int *__restrict__ pts, double *__restrict__ var, double *__restrict__ rhs,
                                const double *__restrict__ pend) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int i = pts[3 * p], j = pts[3 * p + 1], k = pts[3 * p + 2];
  var[offF(g, i, j, k)] = 0.;
  rhs[offT(g, i, j, k)] = pend ? pend[k] : 0.;
}
// solid() with mask on scalar-halo arrays (:772-822); blockIdx.y = scalar field
__global__ void k_ibm_solid_scalar(Geo g, int n, const int *__restrict__ pts, const double *__restrict__ mk, double *__restrict__ var,
                                   long long ssl, double *__restrict__ rhs, long long tsl, double val_, const double *__restrict__ valp = nullptr) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  var += blockIdx.y * ssl;
  rhs += blockIdx.y * tsl;
  const double val = valp ? *valp : val_;   // temperature: the value lives on the device (k_thermo_final)
  const double eps1 = 1.e-10;
  const int i = pts[3 * p], j = pts[3 * p + 1], k = pts[3 * p + 2];
  double v = val, r = 0., count = 0.;
  const int nb[6][3] = {{0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}, {1, 0, 0}, {-1, 0, 0}};
#pragma unroll
  for (int d = 0; d < 6; d++) {
    const int ii = i + nb[d][0], jj = j + nb[d][1], kk = k + nb[d][2];
    if (fabs(mk[offF(g, ii, jj, kk)] - 1.) < eps1) {
      count = count + 1.;
      v = v + var[offS(g, ii, jj, kk)];
      r = r + rhs[offST(g, ii, jj, kk)];
    }
  }
  if (count > 0.) { v = (v - val) / count; r = r / count; }
  var[offS(g, i, j, k)] = v;
  rhs[offST(g, i, j, k)] = r;
}

// diffu_corr / diffv_corr / diffw_corr (:990-1125); COMP 0,1,2
template <int COMP>
__global__ void k_ibm_diffcorr_mom(Geo g, int n, const int *__restrict__ pts, const double *__restrict__ mk,
                                   const double *__restrict__ ekm, const double *__restrict__ vel, double *__restrict__ tend) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double eps1 = 1.e-10;
  const int i = pts[3 * p], j = pts[3 * p + 1], k = pts[3 * p + 2];
#define E(a, b, c) ekm[offF(g, a, b, c)]
#define V(a, b, c) vel[offF(g, a, b, c)]
#define MK(a, b, c) (fabs(mk[offF(g, a, b, c)]) < eps1)
  double t = tend[offT(g, i, j, k)];
  if (COMP == 0) {
    if (MK(i, j + 1, k)) {
      const double empo = 0.25 * ((E(i, j, k) + E(i, j + 1, k)) + (E(i - 1, j, k) + E(i - 1, j + 1, k)));
      t = t - empo * (V(i, j + 1, k) - V(i, j, k)) * g.dy2i;
    }
    if (MK(i, j - 1, k)) {
      const double emmo = 0.25 * ((E(i, j, k) + E(i, j - 1, k)) + (E(i - 1, j - 1, k) + E(i - 1, j, k)));
      t = t + emmo * (V(i, j, k) - V(i, j - 1, k)) * g.dy2i;
    }
    if (MK(i, j, k + 1)) {
      const double emop = (g.dzf[k + 1] * (E(i, j, k) + E(i - 1, j, k)) + g.dzf[k] * (E(i, j, k + 1) + E(i - 1, j, k + 1))) * g.dzhiq[k + 1];
      t = t - emop * (V(i, j, k + 1) - V(i, j, k)) * g.dzhi[k + 1] * g.dzfi[k];
    }
    if (MK(i, j, k - 1)) {
      const double emom = (g.dzf[k - 1] * (E(i, j, k) + E(i - 1, j, k)) + g.dzf[k] * (E(i, j, k - 1) + E(i - 1, j, k - 1))) * g.dzhiq[k];
      t = t + emom * (V(i, j, k) - V(i, j, k - 1)) * g.dzhi[k] * g.dzfi[k];
    }
  } else if (COMP == 1) {
    if (MK(i + 1, j, k)) {
      const double epmo = 0.25 * (E(i, j, k) + E(i, j - 1, k) + E(i + 1, j - 1, k) + E(i + 1, j, k));
      t = t - epmo * (V(i + 1, j, k) - V(i, j, k)) * g.dx2i;
    }
    if (MK(i - 1, j, k)) {
      const double emmo = 0.25 * (E(i, j, k) + E(i, j - 1, k) + E(i - 1, j - 1, k) + E(i - 1, j, k));
      t = t + emmo * (V(i, j, k) - V(i - 1, j, k)) * g.dx2i;
    }
    if (MK(i, j, k + 1)) {
      const double eomp = (g.dzf[k + 1] * (E(i, j, k) + E(i, j - 1, k)) + g.dzf[k] * (E(i, j, k + 1) + E(i, j - 1, k + 1))) * g.dzhiq[k + 1];
      t = t - eomp * (V(i, j, k + 1) - V(i, j, k)) * g.dzhi[k + 1] * g.dzfi[k];
    }
    if (MK(i, j, k - 1)) {
      const double eomm = (g.dzf[k - 1] * (E(i, j, k) + E(i, j - 1, k)) + g.dzf[k] * (E(i, j, k - 1) + E(i, j - 1, k - 1))) * g.dzhiq[k];
      t = t + eomm * (V(i, j, k) - V(i, j, k - 1)) * g.dzhi[k] * g.dzfi[k];
    }
  } else {
    if (MK(i + 1, j, k)) {
      const double epom = (g.dzf[k - 1] * (E(i, j, k) + E(i + 1, j, k)) + g.dzf[k] * (E(i, j, k - 1) + E(i + 1, j, k - 1))) * g.dzhiq[k];
      t = t - epom * (V(i + 1, j, k) - V(i, j, k)) * g.dx2i;
    }
    if (MK(i - 1, j, k)) {
      const double emom = (g.dzf[k - 1] * (E(i, j, k) + E(i - 1, j, k)) + g.dzf[k] * (E(i, j, k - 1) + E(i - 1, j, k - 1))) * g.dzhiq[k];
      t = t + emom * (V(i, j, k) - V(i - 1, j, k)) * g.dx2i;
    }
    if (MK(i, j + 1, k)) {
      const double eopm = (g.dzf[k - 1] * (E(i, j, k) + E(i, j + 1, k)) + g.dzf[k] * (E(i, j, k - 1) + E(i, j + 1, k - 1))) * g.dzhiq[k];
      t = t - eopm * (V(i, j + 1, k) - V(i, j, k)) * g.dy2i;
    }
    if (MK(i, j - 1, k)) {
      const double eomm = (g.dzf[k - 1] * (E(i, j, k) + E(i, j - 1, k)) + g.dzf[k] * (E(i, j, k - 1) + E(i, j - 1, k - 1))) * g.dzhiq[k];
      t = t + eomm * (V(i, j, k) - V(i, j - 1, k)) * g.dy2i;
    }
  }
  tend[offT(g, i, j, k)] = t;
#undef E
#undef V
#undef MK
}

// diffc_corr (:1127-1164); blockIdx.y = scalar field
__global__ void k_ibm_diffcorr_c(Geo g, int n, const int *__restrict__ pts, const double *__restrict__ mk, const double *__restrict__ ekh,
                                 const double *__restrict__ var, long long ssl, double *__restrict__ rhs, long long tsl) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  var += blockIdx.y * ssl;
  rhs += blockIdx.y * tsl;
  const double eps1 = 1.e-10;
  const int i = pts[3 * p], j = pts[3 * p + 1], k = pts[3 * p + 2];
#define E(a, b, c) ekh[offF(g, a, b, c)]
#define S_(a, b, c) var[offS(g, a, b, c)]
#define MK(a, b, c) (fabs(mk[offF(g, a, b, c)]) < eps1)
  double t = rhs[offST(g, i, j, k)];
  if (MK(i + 1, j, k)) t = t - 0.5 * (E(i + 1, j, k) + E(i, j, k)) * (S_(i + 1, j, k) - S_(i, j, k)) * g.dx2i;
  if (MK(i - 1, j, k)) t = t + 0.5 * (E(i, j, k) + E(i - 1, j, k)) * (S_(i, j, k) - S_(i - 1, j, k)) * g.dx2i;
  if (MK(i, j + 1, k)) t = t - 0.5 * (E(i, j + 1, k) + E(i, j, k)) * (S_(i, j + 1, k) - S_(i, j, k)) * g.dy2i;
  if (MK(i, j - 1, k)) t = t + 0.5 * (E(i, j, k) + E(i, j - 1, k)) * (S_(i, j, k) - S_(i, j - 1, k)) * g.dy2i;
  if (MK(i, j, k + 1))
    t = t - 0.5 * (g.dzf[k + 1] * E(i, j, k) + g.dzf[k] * E(i, j, k + 1)) * (S_(i, j, k + 1) - S_(i, j, k)) * g.dzh2i[k + 1] * g.dzfi[k];
  if (MK(i, j, k - 1))
    t = t + 0.5 * (g.dzf[k - 1] * E(i, j, k) + g.dzf[k] * E(i, j, k - 1)) * (S_(i, j, k) - S_(i, j, k - 1)) * g.dzh2i[k] * g.dzfi[k];
  rhs[offST(g, i, j, k)] = t;
#undef E
#undef S_
#undef MK
}

}  // namespace udg
