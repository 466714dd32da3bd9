// scalar_v1.cuh — passive scalars (sv): advecc_kappa / advecc_2nd (src/modadvection.f90:316-421,
// :103-155), diffc (src/modsubgrid.f90:540-623), their part of tstep_integrate
// (src/modtstep.f90:216-218,327,336) and of boundary (fluxtopscal, src/modboundary.f90:1521-1537).
// Scalar arrays carry halo ihc (2 with kappa, src/modglobal.f90:602-609); u0,v0,w0,ekh carry halo 1.
// One thread per cell; every face value of the kappa scheme is evaluated by both cells that share
// the face (bitwise identical), which keeps the kernel free of temporaries (the reference allocates
// two full 3-D arrays and makes three whole-array passes per direction, :327-328,355,379,404).
#pragma once
#include "common.cuh"

namespace udg {

// rlim: src/modadvection.f90:408-421
__device__ __forceinline__ double rlim(double d1, double d2) {
  const double eps1 = 1.e-10;
  const double ri = (d2 + eps1) / (d1 + eps1);
  const double phir = fmax(0., fmin(2. * ri, fmin(1. / 3. + 2. / 3. * ri, 2.)));
  return 0.5 * phir * d1;
}

// face value cf of the kappa scheme at the lower face of cell index `c` along a direction with
// element stride st: vel = advecting velocity at that face; hci_m1, hci_0, hci_p1 = inverse centre
// spacings at c-1, c, c+1; fc = cell size used to scale the limiter (src/modadvection.f90:335-351).
__device__ __forceinline__ double kappa_face(const double *__restrict__ v, long long c, long long st, double vel,
                                             double hci_m1, double hci_0, double hci_p1, double fc) {
  double d1, d2, cf;
  if (vel > 0) {
    d1 = (v[c - st] - v[c - 2 * st]) * hci_m1;
    d2 = (v[c] - v[c - st]) * hci_0;
    cf = v[c - st];
  } else {
    d1 = (v[c] - v[c + st]) * hci_p1;
    d2 = (v[c - st] - v[c]) * hci_0;
    cf = v[c];
  }
  return cf + fc * rlim(d1, d2);
}

// SCHEME 7 = kappa, 2 = cd2.  ADV / DIFF select the operators (advection() / subgrid()).
template <int SCHEME, bool ADV, bool DIFF, bool ACC, bool LES>
__global__ void __launch_bounds__(256) k_scalar_tend(Geo g, const double *__restrict__ u0, const double *__restrict__ v0,
                                                     const double *__restrict__ w0, const double *__restrict__ ekh,
                                                     const double *__restrict__ sv, double *__restrict__ svp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  if (i > g.imax || j > g.jmax) return;
  const long long c = offS(g, i, j, k), t = offST(g, i, j, k), m = offF(g, i, j, k);
  const long long sj = g.pic, sk = g.pkc, mj = g.pi, mk = g.pk;
  double r = ACC ? svp[t] : 0.0;
  if (ADV) {
    if (SCHEME == 7) {
      const double dxi = g.dxi, dx = g.dx, dyi = g.dyi;
      // x (uniform: dxhci = dxi, dxfc = dx, dxfci = dxi)
      const double ul = u0[m], ur = u0[m + 1];
      const double cl = kappa_face(sv, c, 1, ul, dxi, dxi, dxi, dx);
      const double cr = kappa_face(sv, c + 1, 1, ur, dxi, dxi, dxi, dx);
      r = r + (-cr * ur * dxi) + cl * ul * dxi;          // varp + dumu + duml  (:355)
      // y (no stretching: unit spacings, limiter unscaled)
      const double vl = v0[m], vr = v0[m + mj];
      const double cs = kappa_face(sv, c, sj, vl, 1., 1., 1., 1.);
      const double cn = kappa_face(sv, c + sj, sj, vr, 1., 1., 1., 1.);
      r = r + (-cn * vr * dyi) + cs * vl * dyi;          // (:379)
      // z: faces kb+1 .. ke+1 only (:383)
      const double wt = w0[m + mk];
      const double ct = kappa_face(sv, c + sk, sk, wt, g.dzhci[k], g.dzhci[k + 1], g.dzhci[k + 2], g.dzfc[k + 1]);
      double dl = 0.0;
      if (k >= 2) {
        const double wb = w0[m];
        const double cb = kappa_face(sv, c, sk, wb, g.dzhci[k - 1], g.dzhci[k], g.dzhci[k + 1], g.dzfc[k]);
        dl = cb * wb * g.dzfci[k];
      }
      r = r + (-ct * wt * g.dzfci[k]) + dl;              // (:404)
    } else {
      const double dzfk = g.dzf[k], dzfkp = g.dzf[k + 1], dzfkm = g.dzf[k - 1];
      r = r - ((u0[m + 1] * (sv[c + 1] + sv[c]) - u0[m] * (sv[c - 1] + sv[c])) * g.dxi5 +
               (v0[m + mj] * (sv[c + sj] + sv[c]) - v0[m] * (sv[c - sj] + sv[c])) * g.dyi5);
      r = r - (w0[m + mk] * (sv[c + sk] * dzfk + sv[c] * dzfkp) * g.dzhi[k + 1] -
               w0[m] * (sv[c - sk] * dzfk + sv[c] * dzfkm) * g.dzhi[k]) * g.dzfi5[k];
    }
  }
  if (DIFF) {
    if (LES) {
      const double dzfk = g.dzf[k], dzfkp = g.dzf[k + 1], dzfkm = g.dzf[k - 1];
      r = r + 0.5 * (((ekh[m + 1] + ekh[m]) * (sv[c + 1] - sv[c]) - (ekh[m] + ekh[m - 1]) * (sv[c] - sv[c - 1])) * g.dx2i +
                     ((ekh[m + mj] + ekh[m]) * (sv[c + sj] - sv[c]) - (ekh[m] + ekh[m - mj]) * (sv[c] - sv[c - sj])) * g.dy2i +
                     ((dzfkp * ekh[m] + dzfk * ekh[m + mk]) * (sv[c + sk] - sv[c]) * g.dzh2i[k + 1] -
                      (dzfkm * ekh[m] + dzfk * ekh[m - mk]) * (sv[c] - sv[c - sk]) * g.dzh2i[k]) * g.dzfi[k]);
    } else {
      const double cekh = g.numol * g.prandtlmoli;
      r = r + ((cekh * (sv[c + 1] - sv[c]) - cekh * (sv[c] - sv[c - 1])) * g.dx2i +
               (cekh * (sv[c + sj] - sv[c]) - cekh * (sv[c] - sv[c - sj])) * g.dy2i +
               (cekh * (sv[c + sk] - sv[c]) * g.dzhi[k + 1] - cekh * (sv[c] - sv[c - sk]) * g.dzhi[k]) * g.dzfi[k]);
    }
  }
  svp[t] = r;
}

// K3: the same operators for NS scalar fields in ONE pass (config 5 "fused multi-field stencil"): the advecting
// velocities (u, v, w at the six faces) and the seven ekh values are loaded once per cell and reused by all
// fields, so n fields cost (4 + 2n) * 8 B/cell instead of 6n * 8 B/cell.  Field n lives at sv + n * ssl
// (svp + n * tsl).  Arithmetic per field is exactly k_scalar_tend's.
template <int SCHEME, bool ADV, bool DIFF, bool ACC, bool LES, int NS>
__global__ void __launch_bounds__(256) k_scalar_tend_multi(Geo g, const double *__restrict__ u0, const double *__restrict__ v0,
                                                           const double *__restrict__ w0, const double *__restrict__ ekh,
                                                           const double *__restrict__ sv, long long ssl, double *__restrict__ svp,
                                                           long long tsl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  if (i > g.imax || j > g.jmax) return;
  const long long c = offS(g, i, j, k), t = offST(g, i, j, k), m = offF(g, i, j, k);
  const long long sj = g.pic, sk = g.pkc, mj = g.pi, mk = g.pk;
  double ul = 0, ur = 0, vl = 0, vr = 0, wb = 0, wt = 0;
  if (ADV) { ul = u0[m]; ur = u0[m + 1]; vl = v0[m]; vr = v0[m + mj]; wb = w0[m]; wt = w0[m + mk]; }
  double e0 = 0, exm = 0, exp_ = 0, eym = 0, eyp = 0, ezm = 0, ezp = 0;
  if (DIFF && LES) {
    e0 = ekh[m]; exm = ekh[m - 1]; exp_ = ekh[m + 1]; eym = ekh[m - mj]; eyp = ekh[m + mj]; ezm = ekh[m - mk]; ezp = ekh[m + mk];
  }
  const double dzfk = g.dzf[k], dzfkp = g.dzf[k + 1], dzfkm = g.dzf[k - 1];
  const double hc_m1 = g.dzhci[k - 1], hc_0 = g.dzhci[k], hc_p1 = g.dzhci[k + 1], hc_p2 = g.dzhci[k + 2];
  const double fc_0 = g.dzfc[k], fc_p1 = g.dzfc[k + 1], dzfci = g.dzfci[k];
#pragma unroll
  for (int n = 0; n < NS; n++) {
    const double *__restrict__ s = sv + n * ssl;
    double r = ACC ? svp[t + n * tsl] : 0.0;
    if (ADV) {
      if (SCHEME == 7) {
        const double dxi = g.dxi, dx = g.dx, dyi = g.dyi;
        const double cl = kappa_face(s, c, 1, ul, dxi, dxi, dxi, dx);
        const double cr = kappa_face(s, c + 1, 1, ur, dxi, dxi, dxi, dx);
        r = r + (-cr * ur * dxi) + cl * ul * dxi;
        const double cs = kappa_face(s, c, sj, vl, 1., 1., 1., 1.);
        const double cn = kappa_face(s, c + sj, sj, vr, 1., 1., 1., 1.);
        r = r + (-cn * vr * dyi) + cs * vl * dyi;
        const double ct = kappa_face(s, c + sk, sk, wt, hc_0, hc_p1, hc_p2, fc_p1);
        double dl = 0.0;
        if (k >= 2) {
          const double cb = kappa_face(s, c, sk, wb, hc_m1, hc_0, hc_p1, fc_0);
          dl = cb * wb * dzfci;
        }
        r = r + (-ct * wt * dzfci) + dl;
      } else {
        r = r - ((ur * (s[c + 1] + s[c]) - ul * (s[c - 1] + s[c])) * g.dxi5 +
                 (vr * (s[c + sj] + s[c]) - vl * (s[c - sj] + s[c])) * g.dyi5);
        r = r - (wt * (s[c + sk] * dzfk + s[c] * dzfkp) * g.dzhi[k + 1] -
                 wb * (s[c - sk] * dzfk + s[c] * dzfkm) * g.dzhi[k]) * g.dzfi5[k];
      }
    }
    if (DIFF) {
      if (LES) {
        r = r + 0.5 * (((exp_ + e0) * (s[c + 1] - s[c]) - (e0 + exm) * (s[c] - s[c - 1])) * g.dx2i +
                       ((eyp + e0) * (s[c + sj] - s[c]) - (e0 + eym) * (s[c] - s[c - sj])) * g.dy2i +
                       ((dzfkp * e0 + dzfk * ezp) * (s[c + sk] - s[c]) * g.dzh2i[k + 1] -
                        (dzfkm * e0 + dzfk * ezm) * (s[c] - s[c - sk]) * g.dzh2i[k]) * g.dzfi[k]);
      } else {
        const double cekh = g.numol * g.prandtlmoli;
        r = r + ((cekh * (s[c + 1] - s[c]) - cekh * (s[c] - s[c - 1])) * g.dx2i +
                 (cekh * (s[c + sj] - s[c]) - cekh * (s[c] - s[c - sj])) * g.dy2i +
                 (cekh * (s[c + sk] - s[c]) * g.dzhi[k + 1] - cekh * (s[c] - s[c - sk]) * g.dzhi[k]) * g.dzfi[k]);
      }
    }
    svp[t + n * tsl] = r;
  }
}

// branch-free form of kappa_face on register operands: the face lies between vm (cell c-1) and v0 (cell c)
__device__ __forceinline__ double kface(double vmm, double vm, double v0, double vp, double vel, double hci_m1, double hci_0,
                                        double hci_p1, double fc) {
  const bool pos = vel > 0;
  const double d1 = pos ? (vm - vmm) * hci_m1 : (v0 - vp) * hci_p1;
  const double d2 = pos ? (v0 - vm) * hci_0 : (vm - v0) * hci_0;
  const double cf = pos ? vm : v0;
  return cf + fc * rlim(d1, d2);
}

// K3, k-marching form for the kappa scheme (advecc_kappa + diffc for NS fields in one pass).  A thread owns an (i,j)
// column over KC levels.  Per field it carries the four levels k-2..k+1 of its column in registers and loads one new
// value per level; the top-face value of level k is the bottom-face value of level k+1 (identical expression,
// src/modadvection.f90:383-404), x-neighbours and the right-face value come from the neighbouring lanes by warp
// shuffle (a warp covers 31 cells + 1 helper lane that only evaluates the face it shares with lane 30).  So a cell
// costs 4 limiter evaluations and ~6 loads per field instead of 6 evaluations (each with its own divergent branch)
// and ~21 loads.  Faces, limiter and accumulation order are the reference's: results equal k_scalar_tend's bit for bit.
constexpr int SC_WX = 31, SC_BY = 8, SC_KC = 32;
template <bool DIFF, bool ACC, bool LES, int NS>
__global__ void __launch_bounds__(32 * SC_BY, 2) k_scalar_kappa_march(Geo g, const double *__restrict__ u0, const double *__restrict__ v0,
                                                                   const double *__restrict__ w0, const double *__restrict__ ekh,
                                                                   const double *__restrict__ sv, long long ssl, double *__restrict__ svp,
                                                                   long long tsl, int pf) {
  const int lane = threadIdx.x;
  const int i = blockIdx.x * SC_WX + lane + 1;
  const int j = blockIdx.y * SC_BY + threadIdx.y + 1;
  const int k0 = blockIdx.z * SC_KC + 1, k1 = min(k0 + SC_KC, g.ktot + 1);
  // helper lane / cells right of the slab: still take part in the shuffles, never store.  Clamp the column so
  // every address stays inside the arrays (halo width 2 gives i <= imax+1 room; j likewise clamped)
  const bool own = (lane < SC_WX) && (i <= g.imax) && (j <= g.jmax);
  // column clamps: the scalar arrays have halo 2, so the lane right of the helper (i = imax+2) still supplies a valid
  // neighbour value; the momentum-halo arrays (halo 1) are clamped one column earlier (their values at a clamped lane
  // are never consumed by an owned cell)
  const int ic = min(i, g.imax + 1), ics = min(i, g.imax + 2), jc = min(j, g.jmax);
  const long long sj = g.pic, sk = g.pkc, mj = g.pi, mk = g.pk;
  const double *ps = sv + offS(g, ics, jc, k0);
  long long m = offF(g, ic, jc, k0);
  long long t = offST(g, ic, jc, k0);
  const double dxi = g.dxi, dx = g.dx, dyi = g.dyi;
  const bool lo0 = lane >= 1, lo1 = lane >= 2, hi0 = lane <= 30;
  double sm2[NS], sm1[NS], s0[NS], sp1[NS], cz[NS];
#pragma unroll
  for (int n = 0; n < NS; n++) {
    const double *q = ps + n * ssl;
    sm2[n] = q[-2 * sk]; sm1[n] = q[-sk]; s0[n] = q[0]; sp1[n] = q[sk];
  }
  double wb = w0[m];
  double e0 = 0, ezm = 0;
  if (DIFF && LES) { e0 = ekh[m]; ezm = ekh[m - mk]; }
#pragma unroll
  for (int n = 0; n < NS; n++)
    cz[n] = (k0 >= 2) ? kface(sm2[n], sm1[n], s0[n], sp1[n], wb, g.dzhci[k0 - 1], g.dzhci[k0], g.dzhci[k0 + 1], g.dzfc[k0]) : 0.0;
  for (int k = k0; k < k1; k++) {
    if (pf > 0 && k + pf <= g.ktot) {   // planes pf levels ahead -> L2 (the kernel is latency-bound)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(u0 + m + pf * mk));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(v0 + m + pf * mk));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(w0 + m + (pf + 1) * mk));
      if (DIFF && LES) asm volatile("prefetch.global.L2 [%0];" ::"l"(ekh + m + (pf + 1) * mk));
#pragma unroll
      for (int n = 0; n < NS; n++) asm volatile("prefetch.global.L2 [%0];" ::"l"(ps + n * ssl + (pf + 2) * sk));
    }
    const double ul = u0[m];
    const double ur = __shfl_down_sync(0xffffffffu, ul, 1);
    const double vl = v0[m], vr = v0[m + mj];
    const double wt = w0[m + mk];
    double exm = 0, exp_ = 0, eym = 0, eyp = 0, ezp = 0;
    if (DIFF && LES) { exm = ekh[m - 1]; exp_ = ekh[m + 1]; eym = ekh[m - mj]; eyp = ekh[m + mj]; ezp = ekh[m + mk]; }
    const double dzfk = g.dzf[k], dzfkp = g.dzf[k + 1], dzfkm = g.dzf[k - 1];
    const double hc_0 = g.dzhci[k], hc_p1 = g.dzhci[k + 1], hc_p2 = g.dzhci[k + 2];
    const double fc_p1 = g.dzfc[k + 1], dzfci = g.dzfci[k];
#pragma unroll
    for (int n = 0; n < NS; n++) {
      const double *q = ps + n * ssl;
      const double sp2 = q[2 * sk];
      // x neighbours of level k: from the neighbouring lanes, edge lanes from memory
      double xm1 = __shfl_up_sync(0xffffffffu, s0[n], 1), xm2 = __shfl_up_sync(0xffffffffu, s0[n], 2);
      double xp1 = __shfl_down_sync(0xffffffffu, s0[n], 1);
      if (!lo0) xm1 = q[-1];
      if (!lo1) xm2 = q[-2];
      if (!hi0) xp1 = (i <= g.imax + 1) ? q[1] : 0.0;   // lane 31: from memory unless it sits beyond the halo
      const double ym1 = q[-sj], ym2 = q[-2 * sj], yp1 = q[sj], yp2 = q[2 * sj];
      const double c = s0[n];
      const double cl = kface(xm2, xm1, c, xp1, ul, dxi, dxi, dxi, dx);
      const double cr = __shfl_down_sync(0xffffffffu, cl, 1);
      double r = ACC ? svp[t + n * tsl] : 0.0;
      r = r + (-cr * ur * dxi) + cl * ul * dxi;
      const double cs = kface(ym2, ym1, c, yp1, vl, 1., 1., 1., 1.);
      const double cn = kface(ym1, c, yp1, yp2, vr, 1., 1., 1., 1.);
      r = r + (-cn * vr * dyi) + cs * vl * dyi;
      const double ct = kface(sm1[n], c, sp1[n], sp2, wt, hc_0, hc_p1, hc_p2, fc_p1);
      const double dl = (k >= 2) ? cz[n] * wb * dzfci : 0.0;
      r = r + (-ct * wt * dzfci) + dl;
      if (DIFF) {
        if (LES) {
          r = r + 0.5 * (((exp_ + e0) * (xp1 - c) - (e0 + exm) * (c - xm1)) * g.dx2i +
                         ((eyp + e0) * (yp1 - c) - (e0 + eym) * (c - ym1)) * g.dy2i +
                         ((dzfkp * e0 + dzfk * ezp) * (sp1[n] - c) * g.dzh2i[k + 1] -
                          (dzfkm * e0 + dzfk * ezm) * (c - sm1[n]) * g.dzh2i[k]) * g.dzfi[k]);
        } else {
          const double cekh = g.numol * g.prandtlmoli;
          r = r + ((cekh * (xp1 - c) - cekh * (c - xm1)) * g.dx2i +
                   (cekh * (yp1 - c) - cekh * (c - ym1)) * g.dy2i +
                   (cekh * (sp1[n] - c) * g.dzhi[k + 1] - cekh * (c - sm1[n]) * g.dzhi[k]) * g.dzfi[k]);
        }
      }
      if (own) svp[t + n * tsl] = r;
      cz[n] = ct;
      sm2[n] = sm1[n]; sm1[n] = c; s0[n] = sp1[n]; sp1[n] = sp2;
    }
    wb = wt; ezm = e0; e0 = ezp;
    ps += sk; m += mk; t += sk;
  }
}

// Reference quirk kept for parity (src/modadvection.f90:355,379: `varp = varp + dumu + duml` are WHOLE-array adds and the
// face loops run to ie+1 / je+1): advecc_kappa leaves the one-sided flux of the first / last face in the lateral halo
// cells of the tendency: varp(ib-1) = -cf(ib) u0(ib) dxfci, varp(ie+1) = +cf(ie+1) u0(ie+1) dxfci, same in y.  Nothing
// on the fluid path reads those cells, but ibmnorm's solid() averages the tendencies of the fluid neighbours of a
// solid point and such a neighbour may lie in the halo (src/modibm.f90:781-815).  Only launched when IBM masking is on.
// One thread per halo cell of the four strips; blockIdx.z = field.
template <bool ACC>
__global__ void k_scalar_kappa_halo_flux(Geo g, const double *__restrict__ u0, const double *__restrict__ v0,
                                         const double *__restrict__ sv, long long ssl, double *__restrict__ svp, long long tsl) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x + 1;   // j (x strips) or i (y strips)
  const int k = blockIdx.y + 1;
  const double *s = sv + blockIdx.z * ssl;
  double *sp = svp + blockIdx.z * tsl;
  const long long sj = g.pic;
  if (a <= g.jmax) {   // x strips: cells (0, j, k) and (imax+1, j, k)
    const int j = a;
    {
      const long long c = offS(g, 1, j, k);
      const double ul = u0[offF(g, 1, j, k)];
      const double cf = kface(s[c - 2], s[c - 1], s[c], s[c + 1], ul, g.dxi, g.dxi, g.dxi, g.dx);
      const long long t = offST(g, 0, j, k);
      sp[t] = (ACC ? sp[t] : 0.0) + (-cf * ul * g.dxi);
    }
    {
      const long long c = offS(g, g.imax + 1, j, k);
      const double ul = u0[offF(g, g.imax + 1, j, k)];
      const double cf = kface(s[c - 2], s[c - 1], s[c], s[c + 1], ul, g.dxi, g.dxi, g.dxi, g.dx);
      const long long t = offST(g, g.imax + 1, j, k);
      sp[t] = (ACC ? sp[t] : 0.0) + cf * ul * g.dxi;
    }
  }
  if (a <= g.imax) {   // y strips: cells (i, 0, k) and (i, jmax+1, k)
    const int i = a;
    {
      const long long c = offS(g, i, 1, k);
      const double vl = v0[offF(g, i, 1, k)];
      const double cf = kface(s[c - 2 * sj], s[c - sj], s[c], s[c + sj], vl, 1., 1., 1., 1.);
      const long long t = offST(g, i, 0, k);
      sp[t] = (ACC ? sp[t] : 0.0) + (-cf * vl * g.dyi);
    }
    {
      const long long c = offS(g, i, g.jmax + 1, k);
      const double vl = v0[offF(g, i, g.jmax + 1, k)];
      const double cf = kface(s[c - 2 * sj], s[c - sj], s[c], s[c + sj], vl, 1., 1., 1., 1.);
      const long long t = offST(g, i, g.jmax + 1, k);
      sp[t] = (ACC ? sp[t] : 0.0) + cf * vl * g.dyi;
    }
  }
}

// sv0 = svm + rk3coef*svp ; (step 3) svm = sv0   — src/modtstep.f90:216-218,336
template <bool STEP3>
__global__ void __launch_bounds__(256) k_scalar_integrate(Geo g, double rk3coef, double *__restrict__ sv0, double *__restrict__ svm,
                                                          const double *__restrict__ svp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  if (i > g.imax || j > g.jmax) return;
  const long long c = offS(g, i, j, k), t = offST(g, i, j, k);
  const double a = svm[c] + rk3coef * svp[t];
  sv0[c] = a;
  if (STEP3) svm[c] = a;
}

// fluxtopscal with zero flux (src/modboundary.f90:1521-1537): ghost levels ke+1..ke+khc copy level ke on the
// momentum-halo footprint (ib-ih:ie+ih, jb-jh:je+jh) of the scalar arrays
__global__ void k_scalar_top(Geo g, double *__restrict__ sv0, double *__restrict__ svm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1 - g.ih;
  const int j = blockIdx.y + 1 - g.jh;
  if (i > g.imax + g.ih) return;
  const long long cK = offS(g, i, j, g.ktot);
  for (int mm = 1; mm <= g.khc; mm++) {
    sv0[cK + mm * g.pkc] = sv0[cK];
    svm[cK + mm * g.pkc] = svm[cK];
  }
}

}  // namespace udg
