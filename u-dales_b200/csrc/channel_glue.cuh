// channel_glue.cuh — the rest of the resident-channel glue of examples/999:
//   bottom -> wfmneutral(.., 91)   src/modibm.f90:1998-2100, src/modwallfunctions.f90:307-349 (+ the zero-flux scalar
//                                  bottom correction, src/modibm.f90:2077-2091)
//   masscorr, volume-flow branches src/modforces.f90:394-420 (u), :470-495 (v) with avexy_ibm (src/modmpi.f90:623-664)
#pragma once
#include "common.cuh"

namespace udg {

// wfmneutral case 91 on the plane k = kb: one thread per (i,j) does the u and the v update of its cell.  dxf = dx,
// dxhi = dxi (x uniform).  momfluxb accumulates both contributions like the reference (:326, :343).
__global__ void __launch_bounds__(256) k_bottom_wfmneutral(Geo g, double z0, double fkar, const double *__restrict__ u0, const double *__restrict__ v0,
                                                           const double *__restrict__ ekm, double *__restrict__ up, double *__restrict__ vp,
                                                           double *__restrict__ momfluxb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  if (i > g.imax || j > g.jmax) return;
  const int k = 1, km = 0;
  const long long c = offF(g, i, j, k), t = offT(g, i, j, k), sj = g.pi, sk = g.pk;
  const double fkar2 = fkar * fkar, umin = 0.0001;
  const double dzfk = g.dzf[k], dzfkm = g.dzf[km], dzfik = g.dzfi[k], dzhik = g.dzhi[k], dzhiqk = g.dzhiq[k];
  const double delta = 0.5 * dzfk;
  const double lg = log(delta / z0);
  const double ctm = fkar2 / (lg * lg);
  double mf = momfluxb[c];
  {
    const double ut1 = u0[c];
    const double ut2 = (v0[c] + v0[c - 1] + v0[c + sj] + v0[c - 1 + sj]) * 0.25;
    const double utang = fmax(umin, (ut1 * ut1 + ut2 * ut2));
    const double bcmomflux = copysign(fabs(ut1) * sqrt(utang) * ctm, ut1);
    mf = mf + bcmomflux * dzfik;
    const double emom = (dzfkm * (ekm[c] * g.dx + ekm[c - 1] * g.dx) + dzfk * (ekm[c - sk] * g.dx + ekm[c - 1 - sk] * g.dx)) * g.dxi * dzhiqk;
    up[t] = up[t] + (ut1 - u0[c - sk]) * emom * dzhik * dzfik - bcmomflux * dzfik;
  }
  {
    const double ut1 = (u0[c] + u0[c - sj] + u0[c + 1 - sj] + u0[c + 1]) * 0.25;
    const double ut2 = v0[c];
    const double utang = fmax(umin, (ut1 * ut1 + ut2 * ut2));
    const double bcmomflux = copysign(fabs(ut2) * sqrt(utang) * ctm, ut2);
    mf = mf + bcmomflux * dzfik;
    const double eomm = (dzfkm * (ekm[c] + ekm[c - sj]) + dzfk * (ekm[c - sk] + ekm[c - sj - sk])) * dzhiqk;
    vp[t] = vp[t] + (ut2 - v0[c - sk]) * eomm * dzhik * dzfik - bcmomflux * dzfik;
  }
  momfluxb[c] = mf;
}
// src/modibm.f90:2077-2091 (BCbots = 1, zero surface flux: add = 0.) and :2033-2046 (BCbotT = 1: add = -wtsurf);
// blockIdx.z = scalar index
__global__ void __launch_bounds__(256) k_bottom_scalar(Geo g, const double *__restrict__ ekh, const double *__restrict__ sv0, long long ssl,
                                                       double *__restrict__ svp, long long tsl, double add) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  if (i > g.imax || j > g.jmax) return;
  const int kb = 1;
  const double *s = sv0 + blockIdx.z * ssl;
  double *sp = svp + blockIdx.z * tsl;
  const long long c = offS(g, i, j, kb), t = offST(g, i, j, kb), m = offF(g, i, j, kb);
  sp[t] = sp[t] + (0.5 * (g.dzf[kb - 1] * ekh[m] + g.dzf[kb] * ekh[m - g.pk]) * (s[c] - s[c - g.pkc]) * g.dzh2i[kb] + add) * g.dzfi[kb];
}

// ---- wfuno: wall functions with stability correction (Uno 1995 / Cai 2012), src/modwallfunctions.f90:24-260 ----------
// unom (:226-260) and unoh (:176-223); logdz = log(delta / z0), logzh = log(z0 / z0h), sqdz = sqrt(delta / z0) come from
// the host (the same libm the reference uses)
__device__ __forceinline__ void uno_f(double Ri, double fkar2, double logdz, double sqdz, double &Fm, double &Fh) {
  const double b1 = 9.4, b2 = 4.7, dm = 7.4, dh = 5.3;
  if (Ri > 0.) { Fm = 1. / ((1. + b2 * Ri) * (1. + b2 * Ri)); Fh = Fm; }
  else {
    const double cm = (dm * fkar2) / (logdz * logdz) * b1 * sqdz, ch = (dh * fkar2) / (logdz * logdz) * b1 * sqdz;
    const double sr = sqrt(fabs(Ri));
    Fm = 1. - (b1 * Ri) / (1. + cm * sr);
    Fh = 1. - (b1 * Ri) / (1. + ch * sr);
  }
}
__device__ __forceinline__ double unom(double logdz, double logzh, double sqdz, double Ribl0, double fkar2, double pt) {
  double Fm, Fh;
  uno_f(Ribl0, fkar2, logdz, sqdz, Fm, Fh);
  const double M = pt * logdz * sqrt(Fm) / Fh;
  const double Ribl1 = Ribl0 - Ribl0 * pt * logzh / (pt * logzh + M);
  uno_f(Ribl1, fkar2, logdz, sqdz, Fm, Fh);
  return fkar2 / (logdz * logdz) * Fm;
}
__device__ __forceinline__ double unoh(double logdz, double logzh, double sqdz, double utangInt, double dT, double Ribl0, double fkar2, double pt) {
  double Fm, Fh;
  uno_f(Ribl0, fkar2, logdz, sqdz, Fm, Fh);
  double M = pt * logdz * sqrt(Fm) / Fh;
  const double Ribl1 = Ribl0 - Ribl0 * pt * logzh / (pt * logzh + M);
  uno_f(Ribl1, fkar2, logdz, sqdz, Fm, Fh);
  M = pt * logdz * sqrt(Fm) / Fh;
  const double dTrough = dT * 1. / (pt * logzh / M + 1.);
  const double octh = sqrt(utangInt) * fkar2 / (logdz * logdz) * Fh / pt;
  return octh * dTrough;
}
struct WfunoPar { double fkar, logdz, logzh, sqdz, delta, grav, twall, pt, tcell; };
// case 91 (:79-128): surface momentum flux on the plane k = kb.  thl0 == nullptr: no temperature equation, Tcell = tcell
__global__ void __launch_bounds__(256) k_bottom_wfuno_mom(Geo g, WfunoPar w, const double *__restrict__ u0, const double *__restrict__ v0,
                                                          const double *__restrict__ thl0, const double *__restrict__ ekm, double *__restrict__ up,
                                                          double *__restrict__ vp, double *__restrict__ momfluxb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  if (i > g.imax || j > g.jmax) return;
  const int k = 1, km = 0;
  const long long c = offF(g, i, j, k), t = offT(g, i, j, k), sj = g.pi, sk = g.pk;
  const double fkar2 = w.fkar * w.fkar, umin = 0.0001, Twall = w.twall;
  const double dzfk = g.dzf[k], dzfkm = g.dzf[km], dzfik = g.dzfi[k], dzhik = g.dzhi[k], dzhiqk = g.dzhiq[k];
  const double tc = thl0 ? thl0[c] : w.tcell;
  double mf = momfluxb[c];
  {
    const double ut1 = u0[c];
    const double ut2 = (v0[c] + v0[c - 1] + v0[c + sj] + v0[c - 1 + sj]) * 0.25;
    const double utang = fmax(umin, (ut1 * ut1 + ut2 * ut2));
    const double dT = ((tc + (thl0 ? thl0[c - 1] : w.tcell)) - (Twall + Twall)) * 0.5;
    const double Ribl0 = w.grav * w.delta * dT * 2 / ((Twall + Twall) * utang);
    const double ctm = unom(w.logdz, w.logzh, w.sqdz, Ribl0, fkar2, w.pt);
    const double bcmomflux = copysign(fabs(ut1) * sqrt(utang) * ctm, ut1);
    mf = mf + bcmomflux * dzfik;
    const double emom = (dzfkm * (ekm[c] * g.dx + ekm[c - 1] * g.dx) + dzfk * (ekm[c - sk] * g.dx + ekm[c - 1 - sk] * g.dx)) * g.dxi * dzhiqk;
    up[t] = up[t] + (ut1 - u0[c - sk]) * emom * dzhik * dzfik - bcmomflux * dzfik;
  }
  {
    const double ut1 = (u0[c] + u0[c - sj] + u0[c + 1 - sj] + u0[c + 1]) * 0.25;
    const double ut2 = v0[c];
    const double utang = fmax(umin, (ut1 * ut1 + ut2 * ut2));
    const double dT = ((tc + (thl0 ? thl0[c - sj] : w.tcell)) - (Twall + Twall)) * 0.5;
    const double Ribl0 = w.grav * w.delta * dT * 2 / ((Twall + Twall) * utang);
    const double ctm = unom(w.logdz, w.logzh, w.sqdz, Ribl0, fkar2, w.pt);
    const double bcmomflux = copysign(fabs(ut2) * sqrt(utang) * ctm, ut2);
    mf = mf + bcmomflux * dzfik;
    const double eomm = (dzfkm * (ekm[c] + ekm[c - sj]) + dzfk * (ekm[c - sk] + ekm[c - sj - sk])) * dzhiqk;
    vp[t] = vp[t] + (ut2 - v0[c - sk]) * eomm * dzhik * dzfik - bcmomflux * dzfik;
  }
  momfluxb[c] = mf;
}
// case 92 (:131-161): surface temperature flux for a wall at fixed temperature (BCbotT = 2)
__global__ void __launch_bounds__(256) k_bottom_wfuno_thl(Geo g, WfunoPar w, const double *__restrict__ u0, const double *__restrict__ v0,
                                                          const double *__restrict__ thl0, const double *__restrict__ ekh, double *__restrict__ thlp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  if (i > g.imax || j > g.jmax) return;
  const int k = 1;
  const long long c = offF(g, i, j, k), t = offT(g, i, j, k), sj = g.pi, sk = g.pk;
  const double fkar2 = w.fkar * w.fkar, umin = 0.0001, Twall = w.twall;
  const double ut1 = (u0[c] + u0[c + 1]) * 0.5, ut2 = (v0[c] + v0[c + sj]) * 0.5;
  const double utang = fmax(umin, (ut1 * ut1 + ut2 * ut2));
  const double dT = (thl0[c] - Twall);
  const double Ribl0 = w.grav * w.delta * dT / (Twall * utang);
  const double bcTflux = unoh(w.logdz, w.logzh, w.sqdz, utang, dT, Ribl0, fkar2, w.pt);
  thlp[t] = thlp[t] + 0.5 * (g.dzf[k - 1] * ekh[c] + g.dzf[k] * ekh[c - sk]) * (thl0[c] - thl0[c - sk]) * g.dzh2i[k] * g.dzfi[k] - bcTflux * g.dzfi[k];
}

// ---- masscorr ---------------------------------------------------------------------------------------------------
// Masked plane sums in a fixed order (reproducible run to run): block b of level k sums its strided share of the
// plane, k_masscorr_finish adds the nblk partials in index order.  slot = 2*comp + (0: tendency, 1: m-field).
// mask: momentum-halo shaped real mask (1 fluid / 0 solid; mask_u / mask_v of the IBM path) or nullptr = all fluid.
constexpr int MC_NBLK = 32;
__global__ void __launch_bounds__(256) k_slab_partial(Geo g, const double *__restrict__ tend, const double *__restrict__ mfld,
                                                      const double *__restrict__ mask, int unmask_k1, double *__restrict__ part /* [2][K][MC_NBLK] */) {
  const int k = blockIdx.y + 1, b = blockIdx.x;
  const long long n = (long long)g.imax * g.jmax;
  double s0 = 0., s1 = 0.;
  for (long long q = (long long)b * blockDim.x + threadIdx.x; q < n; q += (long long)MC_NBLK * blockDim.x) {
    const int j = (int)(q / g.imax) + 1, i = (int)(q - (long long)(j - 1) * g.imax) + 1;
    const long long c = offF(g, i, j, k);
    const double m = (mask && !(unmask_k1 && k == 1)) ? mask[c] : 1.0;
    s0 += tend[offT(g, i, j, k)] * m;
    s1 += mfld[c] * m;
  }
  __shared__ double sh0[256], sh1[256];
  sh0[threadIdx.x] = s0; sh1[threadIdx.x] = s1;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if ((int)threadIdx.x < o) { sh0[threadIdx.x] += sh0[threadIdx.x + o]; sh1[threadIdx.x] += sh1[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part[((long long)0 * g.ktot + (k - 1)) * MC_NBLK + b] = sh0[0];
    part[((long long)1 * g.ktot + (k - 1)) * MC_NBLK + b] = sh1[0];
  }
}
// plane sums -> vol[slot][k] (before the cross-rank sum)
__global__ void k_masscorr_reduce(int K, const double *__restrict__ part, double *__restrict__ vol /* [2][K] */) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= 2 * K) return;
  double s = 0.;
  for (int b = 0; b < MC_NBLK; b++) s += part[(long long)q * MC_NBLK + b];
  vol[q] = s;
}
// the scalar part of masscorr for one component (single thread): def = flowrate - (rk3coef <tend> + <m>), volume
// means with dzf weights over zh(ke+1) (src/modforces.f90:408-413).  fpend: a lazily pending forces() has to be part of
// the tendency mean (its profile is uniform in x, y: the masked plane mean of it is the profile itself).
// Writes def and the effective per-level table the fused tderive+integrate kernel subtracts: fe[k] = fpend*f[k] - def/rk3coef.
__global__ void __launch_bounds__(256) k_masscorr_final(int K, const double *__restrict__ vol, const double *__restrict__ cnt /* fluid points per level, K */,
                                 double cnt_ke, int unmask_k1, const double *__restrict__ dzf, double zhtop, double rk3coef, double flowrate,
                                 const double *__restrict__ f /* forcing table, index k */, int fpend, double *__restrict__ def_out,
                                 double *__restrict__ fe) {
  // one block of 256 threads: a thread per level and a fixed shared-memory tree for the two dzf-weighted sums (a single
  // thread walking the levels pays one dependent global load per level: ~90 us at K = 256)
  __shared__ double sh1[256], sh2[256];
  double s1 = 0., s2 = 0.;
  for (int k = threadIdx.x + 1; k <= K; k += blockDim.x) {
    double d = cnt[k - 1];
    if (k == 1 && unmask_k1) d = cnt_ke;                      // src/modmpi.f90:649-652
    double a = d == 0. ? -999. : vol[k - 1] / d, bm = d == 0. ? -999. : vol[K + k - 1] / d;
    if (fpend && d != 0.) a = a - f[k];
    s1 += a * dzf[k];
    s2 += bm * dzf[k];
  }
  sh1[threadIdx.x] = s1; sh2[threadIdx.x] = s2;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if ((int)threadIdx.x < o) { sh1[threadIdx.x] += sh1[threadIdx.x + o]; sh2[threadIdx.x] += sh2[threadIdx.x + o]; }
    __syncthreads();
  }
  const double outflow = rk3coef * sh1[0] / zhtop, old = sh2[0] / zhtop;
  const double def = flowrate - (outflow + old);
  if (threadIdx.x == 0) *def_out = def;
  const double add = def * (1 / rk3coef);
  for (int k = threadIdx.x; k <= K + 1; k += blockDim.x) fe[k] = (fpend ? f[k] : 0.) - add;
}
// up(i,j,k) = up(i,j,k) - fe(k) on the interior of one tendency (eager form of the masscorr / forces shift)
__global__ void __launch_bounds__(256) k_tend_sub_table(Geo g, const double *__restrict__ fe, double *__restrict__ tp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  if (i > g.imax || j > g.jmax) return;
  const long long t = offT(g, i, j, k);
  tp[t] = tp[t] - fe[k];
}
// fluid points per level of a real mask (interior), as doubles (summed across ranks with ncclSum afterwards)
__global__ void k_mask_count(Geo g, const double *__restrict__ mask, double *__restrict__ cnt) {
  const int k = blockIdx.x + 1;
  double s = 0.;
  const long long n = (long long)g.imax * g.jmax;
  for (long long q = threadIdx.x; q < n; q += blockDim.x) {
    const int j = (int)(q / g.imax) + 1, i = (int)(q - (long long)(j - 1) * g.imax) + 1;
    s += mask ? mask[offF(g, i, j, k)] : 1.0;
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) cnt[k - 1] = sh[0];
}

}  // namespace udg
