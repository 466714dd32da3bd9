// stencil_v1.cuh — direct (one thread per cell, operands through L1/L2) forms of the stencil
// routines.  They are the always-available general path (any tile-unfriendly size, any switch
// combination) and the in-library cross-check for the TMA-staged kernels in momtend_tma.cuh.
// Operand order follows the reference expressions so that differences to the oracle are pure
// FMA-contraction rounding.
#pragma once
#include "common.cuh"

namespace udg {

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double sq_(double x) { return x * x; }

// ekm/ekh of interior cell (i,j,k) from the eddy viscosity e (src/modsubgrid.f90:356-360).  With halo != 0 the
// thread also writes everything closurebc (src/modboundary.f90:434-505) derives from this cell: its periodic
// images in y (and in x when unsplit) and, for k = 1 / k = ktot, the bottom / top ghost levels of the cell and
// of its images — so no separate wrap / ghost kernels run afterwards (x-split: the slab exchange follows).
__device__ __forceinline__ void ek_store(const Geo &g, int i, int j, int k, double e, double *__restrict__ ekm,
                                         double *__restrict__ ekh, int halo) {
  const double m = e + g.numol, hh = e * g.prandtli + g.numol * g.prandtlmoli;
  if (!halo) {
    const long long c = offF(g, i, j, k);
    ekm[c] = m; ekh[c] = hh;
    return;
  }
  const double tm = 2. * g.numol, th = 2. * g.numol * g.prandtlmoli;
  auto put1 = [&](double *__restrict__ am, double *__restrict__ ah, int ti, int tj) {
    const long long c = offF(g, ti, tj, k);
    am[c] = m; ah[c] = hh;
    if (k == 1) { am[c - g.pk] = tm - m; ah[c - g.pk] = th - hh; }
    if (k == g.ktot) {
      if (g.BCtopm == 2) { am[c + g.pk] = tm - m; ah[c + g.pk] = th - hh; }
      else { am[c + g.pk] = m; ah[c + g.pk] = hh; }
    }
  };
  auto put = [&](int ti, int tj) { put1(ekm, ekh, ti, tj); };
  const int ix = img_x(g, i), jy = img_y(g, j);
  put(i, j);
  if (ix >= 0) put1(ekm, ekh, ix, j);
  if (jy >= 0) { put(i, jy); if (ix >= 0) put1(ekm, ekh, ix, jy); }
}

// closure: src/modsubgrid.f90:159-412.  MODEL 0 = DNS (:401-404), 1 = Vreman (:269-360),
// 2 = Smagorinsky (:208-267).  Writes interior ekm/ekh including "+ numol" (:263-264,359-360).
// Ghost cells are produced by k_closurebc_*.
template <int MODEL>
__global__ void __launch_bounds__(256) k_closure(Geo g, const double *__restrict__ u0, const double *__restrict__ v0,
                                                 const double *__restrict__ w0, double *__restrict__ ekm,
                                                 double *__restrict__ ekh, int halo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  if (i > g.imax || j > g.jmax) return;
  const long long c = offF(g, i, j, k);
  const long long sj = g.pi, sk = g.pk;
  double e;
  if (MODEL == 0) {
    e = 0.0;
  } else {
    const double dzfk = g.dzf[k], dzfkp = g.dzf[k + 1], dzfkm = g.dzf[k - 1];
    const double dzhik = g.dzhi[k], dzhikp = g.dzhi[k + 1];
#define U(di, dj, dk) __ldg(u0 + c + (di) + (dj)*sj + (dk)*sk)
#define V(di, dj, dk) __ldg(v0 + c + (di) + (dj)*sj + (dk)*sk)
#define W(di, dj, dk) __ldg(w0 + c + (di) + (dj)*sj + (dk)*sk)
    if (MODEL == 1) {
      const double a11 = (U(1, 0, 0) - U(0, 0, 0)) * g.dxi;
      const double a12 = (V(1, 1, 0) + V(1, 0, 0) - V(-1, 1, 0) - V(-1, 0, 0)) * g.dxiq;
      const double a13 = (W(1, 0, 1) + W(1, 0, 0) - W(-1, 0, 1) - W(-1, 0, 0)) * g.dxiq;
      const double a21 = (U(1, 1, 0) + U(0, 1, 0) - U(1, -1, 0) - U(0, -1, 0)) * g.dyiq;
      const double a22 = (V(0, 1, 0) - V(0, 0, 0)) * g.dyi;
      const double a23 = (W(0, 1, 1) + W(0, 1, 0) - W(0, -1, 1) - W(0, -1, 0)) * g.dyiq;
      const double a31 = (((U(1, 0, 1) + U(0, 0, 1)) * dzfk + (U(1, 0, 0) + U(0, 0, 0)) * dzfkp) * dzhikp -
                          ((U(1, 0, 0) + U(0, 0, 0)) * dzfkm + (U(1, 0, -1) + U(0, 0, -1)) * dzfk) * dzhik) *
                         g.dzfiq[k];
      const double a32 = (((V(0, 1, 1) + V(0, 0, 1)) * dzfk + (V(0, 1, 0) + V(0, 0, 0)) * dzfkp) * dzhikp -
                          ((V(0, 1, 0) + V(0, 0, 0)) * dzfkm + (V(0, 1, -1) + V(0, 0, -1)) * dzfk) * dzhik) *
                         g.dzfiq[k];
      const double a33 = (W(0, 0, 1) - W(0, 0, 0)) * g.dzfi[k];
      const double aa = a11 * a11 + a21 * a21 + a31 * a31 + a12 * a12 + a22 * a22 + a32 * a32 + a13 * a13 +
                        a23 * a23 + a33 * a33;
      const double dzf2 = g.dzf2[k];
      const double b11 = g.dx2 * a11 * a11 + g.dy2 * a21 * a21 + dzf2 * a31 * a31;
      const double b22 = g.dx2 * a12 * a12 + g.dy2 * a22 * a22 + dzf2 * a32 * a32;
      const double b12 = g.dx2 * a11 * a12 + g.dy2 * a21 * a22 + dzf2 * a31 * a32;
      const double b33 = g.dx2 * a13 * a13 + g.dy2 * a23 * a23 + dzf2 * a33 * a33;
      const double b13 = g.dx2 * a11 * a13 + g.dy2 * a21 * a23 + dzf2 * a31 * a33;
      const double b23 = g.dx2 * a12 * a13 + g.dy2 * a22 * a23 + dzf2 * a32 * a33;
      const double bb = b11 * b22 - b12 * b12 + b11 * b33 - b13 * b13 + b22 * b33 - b23 * b23;
      e = (bb < 1.e-8) ? 0.0 : g.c_vreman * sqrt(bb / aa);
    } else {
      const double mlen = g.csz * g.delta[k];
      double s2;
#define SQ(x) sq_(x)
      s2 = SQ((U(1, 0, 0) - U(0, 0, 0)) * g.dxi) + SQ((V(0, 1, 0) - V(0, 0, 0)) * g.dyi) +
           SQ((W(0, 0, 1) - W(0, 0, 0)) * g.dzfi[k]);
      s2 = s2 + 0.125 * (SQ((W(0, 0, 1) - W(-1, 0, 1)) * g.dxi + (U(0, 0, 1) - U(0, 0, 0)) * dzhikp) +
                         SQ((W(0, 0, 0) - W(-1, 0, 0)) * g.dxi + (U(0, 0, 0) - U(0, 0, -1)) * dzhik) +
                         SQ((W(1, 0, 0) - W(0, 0, 0)) * g.dxi + (U(1, 0, 0) - U(1, 0, -1)) * dzhik) +
                         SQ((W(1, 0, 1) - W(0, 0, 1)) * g.dxi + (U(1, 0, 1) - U(1, 0, 0)) * dzhikp));
      s2 = s2 + 0.125 * (SQ((U(0, 1, 0) - U(0, 0, 0)) * g.dyi + (V(0, 1, 0) - V(-1, 1, 0)) * g.dxi) +
                         SQ((U(0, 0, 0) - U(0, -1, 0)) * g.dyi + (V(0, 0, 0) - V(-1, 0, 0)) * g.dxi) +
                         SQ((U(1, 0, 0) - U(1, -1, 0)) * g.dyi + (V(1, 0, 0) - V(0, 0, 0)) * g.dxi) +
                         SQ((U(1, 1, 0) - U(1, 0, 0)) * g.dyi + (V(1, 1, 0) - V(0, 1, 0)) * g.dxi));
      s2 = s2 + 0.125 * (SQ((V(0, 0, 1) - V(0, 0, 0)) * dzhikp + (W(0, 0, 1) - W(0, -1, 1)) * g.dyi) +
                         SQ((V(0, 0, 0) - V(0, 0, -1)) * dzhik + (W(0, 0, 0) - W(0, -1, 0)) * g.dyi) +
                         SQ((V(0, 1, 0) - V(0, 1, -1)) * dzhik + (W(0, 1, 0) - W(0, 0, 0)) * g.dyi) +
                         SQ((V(0, 1, 1) - V(0, 1, 0)) * dzhikp + (W(0, 1, 1) - W(0, 0, 1)) * g.dyi));
#undef SQ
      e = (mlen * mlen) * sqrt(2. * s2);
    }
#undef U
#undef V
#undef W
  }
  ek_store(g, i, j, k, e, ekm, ekh, halo);
}

// Vreman closure, k-marching form of k_closure<1>: one thread owns an (i,j) column over KC levels and carries
// everything that is shared between consecutive levels in registers (the level-k and level-(k+1) operands of the
// vertical derivatives, the w ring), so a level costs 17 loads instead of 30 and almost no address arithmetic.
// Operand order inside each gradient is the reference's (src/modsubgrid.f90:271-327): differences to k_closure<1>
// and to the oracle are FMA-contraction rounding only.  64 registers / 4 CTAs per SM is a measured optimum: 3 or 2 CTAs
// with a two-level unrolled loop (more loads in flight per thread) ran 0.28 - 0.31 ms against 0.259 ms, 5 or 6 CTAs
// spill and ran 0.40 - 0.52 ms (profiles/r2_ab6_closure.jsonl).
template <int KC, int PF>
__global__ void __launch_bounds__(256, 4) k_closure_vreman_march(Geo g, const double *__restrict__ u0, const double *__restrict__ v0,
                                                              const double *__restrict__ w0, double *__restrict__ ekm,
                                                              double *__restrict__ ekh, int halo, int rev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  if (i > g.imax || j > g.jmax) return;
  const int cz = rev ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;   // rev: top chunk first (L2 reuse)
  const int k0 = cz * KC + 1, k1 = min(k0 + KC, g.ktot + 1);
  const long long sj = g.pi, sk = g.pk;
  const long long c0 = offF(g, i, j, k0 - 1);
  const double *pu = u0 + c0, *pv = v0 + c0, *pw = w0 + c0;
  const double dxi = g.dxi, dyi = g.dyi, dxiq = g.dxiq, dyiq = g.dyiq, dx2 = g.dx2, dy2 = g.dy2;
  double su_km = __ldg(pu + 1) + __ldg(pu), sv_km = __ldg(pv + sj) + __ldg(pv);   // level k0-1
  pu += sk; pv += sk; pw += sk;                                                   // level k0
  double u_c = __ldg(pu), u_ip = __ldg(pu + 1), v_c = __ldg(pv), v_jp = __ldg(pv + sj);
  double w_c = __ldg(pw), w_ip = __ldg(pw + 1), w_im = __ldg(pw - 1), w_jp = __ldg(pw + sj), w_jm = __ldg(pw - sj);
  for (int k = k0; k < k1; k++) {
    const int K = k + 1;
    if (PF > 0 && k + PF <= g.ktot + 1) {   // pull the planes PF levels ahead into L2 while this level computes: the kernel is
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pu + PF * sk));   // latency-bound (ncu: long-scoreboard stalls, 33 % of DRAM),
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pv + PF * sk));   // 0.259 -> 0.245 ms at 256^3 with PF = 2 (profiles/r2_ab8_closure_pf.jsonl)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pw + PF * sk));
    }
    const double u_ipjp = __ldg(pu + sj + 1), u_jp = __ldg(pu + sj), u_ipjm = __ldg(pu - sj + 1), u_jm = __ldg(pu - sj);
    const double v_ipjp = __ldg(pv + sj + 1), v_ip = __ldg(pv + 1), v_imjp = __ldg(pv + sj - 1), v_im = __ldg(pv - 1);
    const double uK_c = __ldg(pu + sk), uK_ip = __ldg(pu + sk + 1), vK_c = __ldg(pv + sk), vK_jp = __ldg(pv + sk + sj);
    const double wK_c = __ldg(pw + sk), wK_ip = __ldg(pw + sk + 1), wK_im = __ldg(pw + sk - 1);
    const double wK_jp = __ldg(pw + sk + sj), wK_jm = __ldg(pw + sk - sj);
    const double dzfk = __ldg(g.dzf + k), dzfK = __ldg(g.dzf + K), dzfkm = __ldg(g.dzf + k - 1);
    const double dzhik = __ldg(g.dzhi + k), dzhiK = __ldg(g.dzhi + K);
    const double dzfiqk = __ldg(g.dzfiq + k), dzfik = __ldg(g.dzfi + k), dzf2 = __ldg(g.dzf2 + k);
    const double su_k = u_ip + u_c, sv_k = v_jp + v_c;
    const double a11 = (u_ip - u_c) * dxi;
    const double a12 = (v_ipjp + v_ip - v_imjp - v_im) * dxiq;
    const double a13 = (wK_ip + w_ip - wK_im - w_im) * dxiq;
    const double a21 = (u_ipjp + u_jp - u_ipjm - u_jm) * dyiq;
    const double a22 = (v_jp - v_c) * dyi;
    const double a23 = (wK_jp + w_jp - wK_jm - w_jm) * dyiq;
    const double suK = uK_ip + uK_c, svK = vK_jp + vK_c;
    const double a31 = ((suK * dzfk + su_k * dzfK) * dzhiK - (su_k * dzfkm + su_km * dzfk) * dzhik) * dzfiqk;
    const double a32 = ((svK * dzfk + sv_k * dzfK) * dzhiK - (sv_k * dzfkm + sv_km * dzfk) * dzhik) * dzfiqk;
    const double a33 = (wK_c - w_c) * dzfik;
    const double aa = a11 * a11 + a21 * a21 + a31 * a31 + a12 * a12 + a22 * a22 + a32 * a32 + a13 * a13 + a23 * a23 + a33 * a33;
    const double b11 = dx2 * a11 * a11 + dy2 * a21 * a21 + dzf2 * a31 * a31;
    const double b22 = dx2 * a12 * a12 + dy2 * a22 * a22 + dzf2 * a32 * a32;
    const double b12 = dx2 * a11 * a12 + dy2 * a21 * a22 + dzf2 * a31 * a32;
    const double b33 = dx2 * a13 * a13 + dy2 * a23 * a23 + dzf2 * a33 * a33;
    const double b13 = dx2 * a11 * a13 + dy2 * a21 * a23 + dzf2 * a31 * a33;
    const double b23 = dx2 * a12 * a13 + dy2 * a22 * a23 + dzf2 * a32 * a33;
    const double bb = b11 * b22 - b12 * b12 + b11 * b33 - b13 * b13 + b22 * b33 - b23 * b23;
    const double e = (bb < 1.e-8) ? 0.0 : g.c_vreman * sqrt(bb / aa);
    ek_store(g, i, j, k, e, ekm, ekh, halo);
    su_km = su_k; sv_km = sv_k;
    u_c = uK_c; u_ip = uK_ip; v_c = vK_c; v_jp = vK_jp;
    w_c = wK_c; w_ip = wK_ip; w_im = wK_im; w_jp = wK_jp; w_jm = wK_jm;
    pu += sk; pv += sk; pw += sk;
  }
}

// closurebc part 1: src/modboundary.f90:447-465 top/bottom ghost levels on (0..imax+1, 0..jmax+1).
// Run AFTER the lateral halo fill of the interior levels so the ghost levels of the lateral halo
// are consistent (the reference gets the same values through its :476-500 wraps).
__global__ void k_closurebc_topbot(Geo g, double *__restrict__ ekm, double *__restrict__ ekh) {
  const int si = blockIdx.x * blockDim.x + threadIdx.x;
  const int sj = blockIdx.y;
  if (si >= g.pi) return;
  const long long b = (long long)si + (long long)g.pi * sj;
  const long long k1 = b + g.pk * g.kh, kK = b + g.pk * (g.ktot + g.kh - 1);
  const double tm = 2. * g.numol, th = 2. * g.numol * g.prandtlmoli;
  if (g.BCtopm == 2) {
    ekm[kK + g.pk] = tm - ekm[kK];
    ekh[kK + g.pk] = th - ekh[kK];
  } else {
    ekm[kK + g.pk] = ekm[kK];
    ekh[kK + g.pk] = ekh[kK];
  }
  ekm[k1 - g.pk] = tm - ekm[k1];
  ekh[k1 - g.pk] = th - ekh[k1];
}

// ---------------------------------------------------------------------------------------------
// periodic wraps of N momentum-halo arrays (xm_periodic/ym_periodic src/modboundary.f90:508-626,
// closurebc :476-500, bcp :1362-1408).  nlev = number of stored k levels, width = halo width.
struct PtrPack { double *p[8]; int n; };

__global__ void k_wrap_x(PtrPack a, int pi, int pj, int nlev, int imax, int h) {
  // all (j,k) incl. halos; one thread per (j,k,m)
  const long long jk = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (jk >= (long long)pj * nlev) return;
  const long long row = jk * pi;
  for (int f = 0; f < a.n; f++) {
    double *q = a.p[f] + row;
    for (int m = 1; m <= h; m++) {
      q[h - m] = q[h + imax - m];          // (ib-m) = (ie+1-m)
      q[h + imax - 1 + m] = q[h - 1 + m];  // (ie+m) = (ib-1+m)
    }
  }
}
__global__ void k_wrap_y(PtrPack a, int pi, int pj, int nlev, int jmax, int h) {
  const int si = blockIdx.x * blockDim.x + threadIdx.x;
  const int lev = blockIdx.y;
  if (si >= pi) return;
  const long long base = (long long)lev * pi * pj + si;
  for (int f = 0; f < a.n; f++) {
    double *q = a.p[f] + base;
    for (int m = 1; m <= h; m++) {
      q[(long long)(h - m) * pi] = q[(long long)(h + jmax - m) * pi];
      q[(long long)(h + jmax - 1 + m) * pi] = q[(long long)(h - 1 + m) * pi];
    }
  }
}

// x-halo exchange between slabs (replaces exchange_halo_z's E/W phase, 2decomp-fft/src/halo_comm_z.f90:73-110):
// gather the first / last interior column of every (j,k) row incl. halo rows into contiguous send
// buffers, and scatter received columns into the halo columns.  One thread per row.
struct HaloPack { double *f[8]; int nlev[8]; long long off[8]; int n; };
// width-h columns: send buffers hold h doubles per row (m = 0..h-1: interior columns h+m / imax+m)
__global__ void k_halo_pack_x(HaloPack a, int pi, int pj, int imax, int h, double *__restrict__ sendL, double *__restrict__ sendR) {
  const int f = blockIdx.y;
  const long long jk = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (jk >= (long long)pj * a.nlev[f]) return;
  const double *q = a.f[f] + jk * pi;
  for (int m = 0; m < h; m++) {
    sendL[a.off[f] + jk * h + m] = q[h + m];      // first h interior columns  -> left neighbour's right halo
    sendR[a.off[f] + jk * h + m] = q[imax + m];   // last h interior columns   -> right neighbour's left halo
  }
}
__global__ void k_halo_unpack_x(HaloPack a, int pi, int pj, int imax, int h, const double *__restrict__ recvL, const double *__restrict__ recvR) {
  const int f = blockIdx.y;
  const long long jk = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (jk >= (long long)pj * a.nlev[f]) return;
  double *q = a.f[f] + jk * pi;
  for (int m = 0; m < h; m++) {
    q[m] = recvL[a.off[f] + jk * h + m];
    q[h + imax + m] = recvR[a.off[f] + jk * h + m];
  }
}

// The same two kernels with the neighbour rendezvous built in (peer-memory path): no separate barrier launch.
//  * pack + signal: every block stores its lines into the neighbours' windows, fences, and counts itself in; the last
//    block raises this rank's flag (epoch) in both neighbours: "my columns have landed in your window".
//  * wait + unpack: thread 0 of every block spins on the two flags the neighbours raise in MY flag array, then the
//    block scatters the window into the halo columns (window reads bypass L1: the lines were written remotely).
struct HaloSync {
  unsigned long long *flagL, *flagR;    // my slot in the left / right neighbour's flag array
  unsigned long long *mineL, *mineR;    // the slots the left / right neighbour raises in my flag array
  unsigned long long epoch, timeout_ns;
  unsigned *counter;                    // blocks of the pack kernel that have finished (reset by the last one)
  volatile int *status;
};
__device__ __forceinline__ unsigned long long halo_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__global__ void k_halo_pack_signal(HaloPack a, int pi, int pj, int imax, int h, double *__restrict__ sendL, double *__restrict__ sendR, HaloSync hs) {
  const int f = blockIdx.y;
  const long long jk = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (jk < (long long)pj * a.nlev[f]) {
    const double *q = a.f[f] + jk * pi;
    for (int m = 0; m < h; m++) {
      sendL[a.off[f] + jk * h + m] = q[h + m];
      sendR[a.off[f] + jk * h + m] = q[imax + m];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned nblocks = gridDim.x * gridDim.y;
    if (atomicAdd(hs.counter, 1u) + 1u == nblocks) {
      *hs.counter = 0;
      __threadfence_system();
      *(volatile unsigned long long *)hs.flagL = hs.epoch;
      *(volatile unsigned long long *)hs.flagR = hs.epoch;
      __threadfence_system();
    }
  }
}
__global__ void k_halo_wait_unpack(HaloPack a, int pi, int pj, int imax, int h, const double *__restrict__ recvL, const double *__restrict__ recvR, HaloSync hs) {
  if (threadIdx.x == 0) {
    const unsigned long long t0 = halo_now_ns();
    while (*(volatile unsigned long long *)hs.mineL < hs.epoch || *(volatile unsigned long long *)hs.mineR < hs.epoch) {
      if (halo_now_ns() - t0 > hs.timeout_ns) { *hs.status = 1; __threadfence_system(); break; }
    }
  }
  __syncthreads();
  const int f = blockIdx.y;
  const long long jk = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (jk >= (long long)pj * a.nlev[f]) return;
  double *q = a.f[f] + jk * pi;
  for (int m = 0; m < h; m++) {
    q[m] = __ldcg(recvL + a.off[f] + jk * h + m);
    q[h + imax + m] = __ldcg(recvR + a.off[f] + jk * h + m);
  }
}

// ---------------------------------------------------------------------------------------------
// momentum tendencies, direct form.  advecu/v/w_2nd src/modadvection.f90:158-314 and
// diffu/v/w src/modsubgrid.f90:672-997.  ACC: add to the existing tendency (drop-in semantics)
// or overwrite (tendencies are zero on entry, src/modtstep.f90:322-324).
template <bool ADV, bool DIFF, bool ACC, bool LES>
__global__ void __launch_bounds__(256) k_momtend_v1(Geo g, const double *__restrict__ u0, const double *__restrict__ v0,
                                                    const double *__restrict__ w0, const double *__restrict__ pres0,
                                                    const double *__restrict__ ekm, double *__restrict__ up,
                                                    double *__restrict__ vp, double *__restrict__ wp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  if (i > g.imax || j > g.jmax) return;
  const long long c = offF(g, i, j, k);
  const long long t = offT(g, i, j, k);
  const long long sj = g.pi, sk = g.pk;
#define U(di, dj, dk) __ldg(u0 + c + (di) + (dj)*sj + (dk)*sk)
#define V(di, dj, dk) __ldg(v0 + c + (di) + (dj)*sj + (dk)*sk)
#define W(di, dj, dk) __ldg(w0 + c + (di) + (dj)*sj + (dk)*sk)
#define P(di, dj, dk) __ldg(pres0 + c + (di) + (dj)*sj + (dk)*sk)
#define E(di, dj, dk) (LES ? __ldg(ekm + c + (di) + (dj)*sj + (dk)*sk) : g.numol)
  const double dzfk = g.dzf[k], dzfkp = g.dzf[k + 1], dzfkm = g.dzf[k - 1];
  const double dzhik = g.dzhi[k], dzhikp = g.dzhi[k + 1];
  const double dzfik = g.dzfi[k], dzfi5k = g.dzfi5[k];
  const double dxi = g.dxi, dyi = g.dyi;
  double ru = ACC ? up[t] : 0.0, rv = ACC ? vp[t] : 0.0;
  if (ADV) {
    ru = ru - (((U(0, 0, 0) + U(1, 0, 0)) * (U(0, 0, 0) + U(1, 0, 0)) - (U(0, 0, 0) + U(-1, 0, 0)) * (U(0, 0, 0) + U(-1, 0, 0))) * g.dxiq +
               ((U(0, 0, 0) + U(0, 1, 0)) * (V(0, 1, 0) + V(-1, 1, 0)) - (U(0, 0, 0) + U(0, -1, 0)) * (V(0, 0, 0) + V(-1, 0, 0))) * g.dyiq) -
         ((P(0, 0, 0) - P(-1, 0, 0)) * dxi);
    ru = ru - ((U(0, 0, 1) * dzfk + U(0, 0, 0) * dzfkp) * dzhikp * (W(0, 0, 1) + W(-1, 0, 1)) -
               (U(0, 0, 0) * dzfkm + U(0, 0, -1) * dzfk) * dzhik * (W(0, 0, 0) + W(-1, 0, 0))) * 0.5 * dzfi5k;
    rv = rv - (((U(1, 0, 0) + U(1, -1, 0)) * (V(0, 0, 0) + V(1, 0, 0)) - (U(0, 0, 0) + U(0, -1, 0)) * (V(0, 0, 0) + V(-1, 0, 0))) * g.dxiq +
               ((V(0, 1, 0) + V(0, 0, 0)) * (V(0, 0, 0) + V(0, 1, 0)) - (V(0, -1, 0) + V(0, 0, 0)) * (V(0, 0, 0) + V(0, -1, 0))) * g.dyiq) -
         ((P(0, 0, 0) - P(0, -1, 0)) * dyi);
    rv = rv - ((W(0, 0, 1) + W(0, -1, 1)) * (V(0, 0, 1) * dzfk + V(0, 0, 0) * dzfkp) * dzhikp -
               (W(0, 0, 0) + W(0, -1, 0)) * (V(0, 0, -1) * dzfk + V(0, 0, 0) * dzfkm) * dzhik) * 0.5 * dzfi5k;
  }
  if (DIFF) {
    const double dzhiqk = g.dzhiq[k], dzhiqkp = g.dzhiq[k + 1];
    {
      const double emom = LES ? (dzfkm * (E(0, 0, 0) + E(-1, 0, 0)) + dzfk * (E(0, 0, -1) + E(-1, 0, -1))) * dzhiqk : g.numol;
      const double emop = LES ? (dzfkp * (E(0, 0, 0) + E(-1, 0, 0)) + dzfk * (E(0, 0, 1) + E(-1, 0, 1))) * dzhiqkp : g.numol;
      const double empo = LES ? 0.25 * ((E(0, 0, 0) + E(0, 1, 0)) + (E(-1, 0, 0) + E(-1, 1, 0))) : g.numol;
      const double emmo = LES ? 0.25 * ((E(0, 0, 0) + E(0, -1, 0)) + (E(-1, -1, 0) + E(-1, 0, 0))) : g.numol;
      ru = ru + (E(0, 0, 0) * (U(1, 0, 0) - U(0, 0, 0)) - E(-1, 0, 0) * (U(0, 0, 0) - U(-1, 0, 0))) * 2. * g.dx2i +
           (empo * ((U(0, 1, 0) - U(0, 0, 0)) * dyi + (V(0, 1, 0) - V(-1, 1, 0)) * dxi) -
            emmo * ((U(0, 0, 0) - U(0, -1, 0)) * dyi + (V(0, 0, 0) - V(-1, 0, 0)) * dxi)) * dyi +
           (emop * ((U(0, 0, 1) - U(0, 0, 0)) * dzhikp + (W(0, 0, 1) - W(-1, 0, 1)) * dxi) -
            emom * ((U(0, 0, 0) - U(0, 0, -1)) * dzhik + (W(0, 0, 0) - W(-1, 0, 0)) * dxi)) * dzfik;
    }
    {
      const double eomm = LES ? (dzfkm * (E(0, 0, 0) + E(0, -1, 0)) + dzfk * (E(0, 0, -1) + E(0, -1, -1))) * dzhiqk : g.numol;
      const double eomp = LES ? (dzfkp * (E(0, 0, 0) + E(0, -1, 0)) + dzfk * (E(0, 0, 1) + E(0, -1, 1))) * dzhiqkp : g.numol;
      const double emmo = LES ? 0.25 * (E(0, 0, 0) + E(0, -1, 0) + E(-1, -1, 0) + E(-1, 0, 0)) : g.numol;
      const double epmo = LES ? 0.25 * (E(0, 0, 0) + E(0, -1, 0) + E(1, -1, 0) + E(1, 0, 0)) : g.numol;
      rv = rv + (epmo * ((V(1, 0, 0) - V(0, 0, 0)) * dxi + (U(1, 0, 0) - U(1, -1, 0)) * dyi) -
                 emmo * ((V(0, 0, 0) - V(-1, 0, 0)) * dxi + (U(0, 0, 0) - U(0, -1, 0)) * dyi)) * dxi +
           (E(0, 0, 0) * (V(0, 1, 0) - V(0, 0, 0)) - E(0, -1, 0) * (V(0, 0, 0) - V(0, -1, 0))) * 2. * g.dy2i +
           (eomp * ((V(0, 0, 1) - V(0, 0, 0)) * dzhikp + (W(0, 0, 1) - W(0, -1, 1)) * dyi) -
            eomm * ((V(0, 0, 0) - V(0, 0, -1)) * dzhik + (W(0, 0, 0) - W(0, -1, 0)) * dyi)) * dzfik;
    }
  }
  up[t] = ru;
  vp[t] = rv;
  if (k >= 2) {
    double rw = ACC ? wp[t] : 0.0;
    if (ADV) {
      rw = rw - (((W(1, 0, 0) + W(0, 0, 0)) * (dzfkm * U(1, 0, 0) + dzfk * U(1, 0, -1)) -
                  (W(0, 0, 0) + W(-1, 0, 0)) * (dzfkm * U(0, 0, 0) + dzfk * U(0, 0, -1))) * g.dxiq * dzhik +
                 ((W(0, 1, 0) + W(0, 0, 0)) * (dzfkm * V(0, 1, 0) + dzfk * V(0, 1, -1)) -
                  (W(0, 0, 0) + W(0, -1, 0)) * (dzfkm * V(0, 0, 0) + dzfk * V(0, 0, -1))) * g.dyiq * dzhik +
                 ((W(0, 0, 0) + W(0, 0, 1)) * (W(0, 0, 0) + W(0, 0, 1)) - (W(0, 0, 0) + W(0, 0, -1)) * (W(0, 0, 0) + W(0, 0, -1))) * g.dzhiq[k]) -
           ((P(0, 0, 0) - P(0, 0, -1)) * dzhik);
    }
    if (DIFF) {
      const double dzhiqk = g.dzhiq[k];
      const double emom = LES ? (dzfkm * (E(0, 0, 0) + E(-1, 0, 0)) + dzfk * (E(0, 0, -1) + E(-1, 0, -1))) * dzhiqk : g.numol;
      const double eomm = LES ? (dzfkm * (E(0, 0, 0) + E(0, -1, 0)) + dzfk * (E(0, 0, -1) + E(0, -1, -1))) * dzhiqk : g.numol;
      const double eopm = LES ? (dzfkm * (E(0, 0, 0) + E(0, 1, 0)) + dzfk * (E(0, 0, -1) + E(0, 1, -1))) * dzhiqk : g.numol;
      const double epom = LES ? (dzfkm * (E(0, 0, 0) + E(1, 0, 0)) + dzfk * (E(0, 0, -1) + E(1, 0, -1))) * dzhiqk : g.numol;
      rw = rw + (epom * ((W(1, 0, 0) - W(0, 0, 0)) * dxi + (U(1, 0, 0) - U(1, 0, -1)) * dzhik) -
                 emom * ((W(0, 0, 0) - W(-1, 0, 0)) * dxi + (U(0, 0, 0) - U(0, 0, -1)) * dzhik)) * dxi +
           (eopm * ((W(0, 1, 0) - W(0, 0, 0)) * dyi + (V(0, 1, 0) - V(0, 1, -1)) * dzhik) -
            eomm * ((W(0, 0, 0) - W(0, -1, 0)) * dyi + (V(0, 0, 0) - V(0, 0, -1)) * dzhik)) * dyi +
           (E(0, 0, 0) * (W(0, 0, 1) - W(0, 0, 0)) * dzfik - E(0, 0, -1) * (W(0, 0, 0) - W(0, 0, -1)) * g.dzfi[k - 1]) * 2. * dzhik;
    }
    wp[t] = rw;
  } else if (!ACC) {
    wp[t] = 0.0;
  }
#undef U
#undef V
#undef W
#undef P
#undef E
}

// ---------------------------------------------------------------------------------------------
// fillps + bcpup: src/modpois.f90:911-973, src/modboundary.f90:1191-1255,1307-1315.
// pup/pvp/pwp are never materialised: p = d/dx(up+um/c) + ... evaluated on the fly.  The +1
// neighbour in x / y is taken from the periodic image when the direction is unsplit (XWRAP /
// YWRAP), otherwise from the halo cell (filled by the halo exchange of up+um/c's inputs).
// Output: halo-free rhs(imax,jmax,ktot).
template <bool XWRAP, bool YWRAP>
__global__ void __launch_bounds__(256) k_fillps(Geo g, double rk3coefi, const double *__restrict__ up, const double *__restrict__ vp,
                                                const double *__restrict__ wp, const double *__restrict__ um,
                                                const double *__restrict__ vm, const double *__restrict__ wm,
                                                double *__restrict__ rhs, int rev, int koff) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  // levels koff+1 .. koff+gridDim.z; rev: top level first (L2 reuse, see LineDesc::rev)
  const int k = koff + (rev ? (int)(gridDim.z - blockIdx.z) : (int)blockIdx.z + 1);
  if (i > g.imax || j > g.jmax) return;
  const int ip = (XWRAP && i == g.imax) ? 1 : i + 1;
  const int jp = (YWRAP && j == g.jmax) ? 1 : j + 1;
  const double pu0 = up[offT(g, i, j, k)] + um[offF(g, i, j, k)] * rk3coefi;
  const double pu1 = up[offT(g, ip, j, k)] + um[offF(g, ip, j, k)] * rk3coefi;
  const double pv0 = vp[offT(g, i, j, k)] + vm[offF(g, i, j, k)] * rk3coefi;
  const double pv1 = vp[offT(g, i, jp, k)] + vm[offF(g, i, jp, k)] * rk3coefi;
  const double pw0 = (k == 1) ? 0.0 : wp[offT(g, i, j, k)] + wm[offF(g, i, j, k)] * rk3coefi;
  const double pw1 = (k == g.ktot) ? 0.0 : wp[offT(g, i, j, k + 1)] + wm[offF(g, i, j, k + 1)] * rk3coefi;
  rhs[offR(g, i, j, k)] = (pu1 - pu0) * g.dxi + (pv1 - pv0) * g.dyi + (pw1 - pw0) * g.dzfi[k];
}

// forces, neutral branch (src/modforces.f90:88-125): up -= dpdxl(k), vp -= dpdyl(k), wp(kb) = 0.  Only launched when
// somebody looks at the tendencies between forces() and tstep_integrate(); otherwise the subtraction is done inside
// the fused tderive+integrate kernel (fx, fy tables; zero tables when no forcing is pending).
__global__ void __launch_bounds__(256) k_forces(Geo g, const double *__restrict__ fx, const double *__restrict__ fy,
                                                double *__restrict__ up, double *__restrict__ vp, double *__restrict__ wp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  if (i > g.imax || j > g.jmax) return;
  const long long t = offT(g, i, j, k);
  up[t] = up[t] - fx[k];
  vp[t] = vp[t] - fy[k];
  if (k == 1) wp[t] = 0.0;
}

// tderive: src/modpois.f90:1046-1056 (velocity tendencies) on a halo'd p.
__global__ void __launch_bounds__(256) k_tderive(Geo g, const double *__restrict__ p, double *__restrict__ up,
                                                 double *__restrict__ vp, double *__restrict__ wp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  if (i > g.imax || j > g.jmax) return;
  const long long c = offF(g, i, j, k), t = offT(g, i, j, k);
  const double pc = p[c];
  up[t] = up[t] - (pc - p[c - 1]) * g.dxi;
  vp[t] = vp[t] - (pc - p[c - g.pi]) * g.dyi;
  if (k >= 2) wp[t] = wp[t] - (pc - p[c - g.pk]) * g.dzhi[k];
}
// pres0 += p on (ib-1:ie+1, jb-1:je+1, kb-1:ke+1): src/modpois.f90:1096-1102
__global__ void k_pres_update(long long n, const double *__restrict__ p, double *__restrict__ pres0) {
  long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; q < n; q += stride) pres0[q] = pres0[q] + p[q];
}


// tderive + tstep_integrate fused (src/modpois.f90:1046-1056,1096-1102 + src/modtstep.f90:171-340): one pass
// that reads p (halo'd, after bcp), the tendencies and the m-fields and writes u0,v0,w0 and pres0 on
// the interior; the projected tendencies themselves are never written back (they are zero after
// tstep_integrate anyway, src/modtstep.f90:322-324).  STEP3: um = u0 too (:330-338).
template <bool STEP3>
__global__ void __launch_bounds__(256) k_tderive_integrate(Geo g, double rk3coef, const double *__restrict__ p,
                                                           const double *__restrict__ up, const double *__restrict__ vp,
                                                           const double *__restrict__ wp, double *__restrict__ um,
                                                           double *__restrict__ vm, double *__restrict__ wm,
                                                           double *__restrict__ u0, double *__restrict__ v0,
                                                           double *__restrict__ w0, double *__restrict__ pres0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  if (i > g.imax || j > g.jmax) return;
  const long long c = offF(g, i, j, k), t = offT(g, i, j, k);
  const double pc = p[c];
  const double ru = up[t] - (pc - p[c - 1]) * g.dxi;
  const double rv = vp[t] - (pc - p[c - g.pi]) * g.dyi;
  double rw = wp[t];
  if (k >= 2) rw = rw - (pc - p[c - g.pk]) * g.dzhi[k];
  const double a = um[c] + rk3coef * ru;
  const double b = vm[c] + rk3coef * rv;
  const double d = wm[c] + rk3coef * rw;
  u0[c] = a; v0[c] = b; w0[c] = d;
  if (STEP3) { um[c] = a; vm[c] = b; wm[c] = d; }
  pres0[c] = pres0[c] + pc;
}
// Same pass, additionally owning everything `bcp`, `halos` and `boundary` do for the cells it writes (periodic
// x,y; src/modboundary.f90:1362-1408, :508-626, :163-204): p(i-1), p(j-1) are read from the periodic image instead
// of a pre-wrapped halo; every new value is also stored to the cell's periodic images (y always, x when unsplit);
// cells at k = ktot also set the top ghost level (freeslip copy / noslip mirror, w = 0); pres0 += p reaches the
// image cells too (:1096-1102).  Only legal when the m-fields / w(k=1) / tendencies at k=1 are in the state the
// reference's own halos+boundary leave them in (the host tracks that; otherwise the plain kernel + wraps run).
// XS: 0 = x unsplit (periodic images in x written here), 1 = x split over GPUs (halo columns of pres0 updated from the
// exchanged p)
template <bool STEP3, int XS = 0, bool FORCE = false>
__global__ void __launch_bounds__(256) k_tderive_integrate_halo(Geo g, double rk3coef, const double *__restrict__ p,
                                                                const double *__restrict__ up, const double *__restrict__ vp,
                                                                const double *__restrict__ wp, double *__restrict__ um,
                                                                double *__restrict__ vm, double *__restrict__ wm,
                                                                double *__restrict__ u0, double *__restrict__ v0,
                                                                double *__restrict__ w0, double *__restrict__ pres0,
                                                                const double *__restrict__ fx, const double *__restrict__ fy, int fz, int koff) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = koff + blockIdx.z + 1;   // levels koff+1 .. koff+gridDim.z
  if (i > g.imax || j > g.jmax) return;
  const long long c = offF(g, i, j, k), t = offT(g, i, j, k);
  const double pc0 = p[c];
  const long long cim = (g.wrapx && i == 1) ? c + (g.imax - 1) : c - 1;
  const long long cjm = (j == 1) ? c + (long long)g.pi * (g.jmax - 1) : c - g.pi;
  // FORCE: a pending forces() (src/modforces.f90:88-125) is applied here: up - dpdxl(k), vp - dpdyl(k), wp(kb) = 0.
  // (A template switch, not zero tables: the two extra per-level loads cost 30 us on this bandwidth-bound kernel.)
  const double ru = (FORCE ? up[t] - __ldg(fx + k) : up[t]) - (pc0 - p[cim]) * g.dxi;
  const double rv = (FORCE ? vp[t] - __ldg(fy + k) : vp[t]) - (pc0 - p[cjm]) * g.dyi;
  double rw = (FORCE && fz && k == 1) ? 0.0 : wp[t];
  if (k >= 2) rw = rw - (pc0 - p[c - g.pk]) * g.dzhi[k];
  const double a = um[c] + rk3coef * ru;
  const double b = vm[c] + rk3coef * rv;
  const double d = wm[c] + rk3coef * rw;
  const bool top = (k == g.ktot);
  const double at = (g.BCtopm == 2) ? 2. * g.Uinf - a : a, bt = (g.BCtopm == 2) ? 2. * g.Vinf - b : b;
  // velocities (+ top ghost) of one cell image into one set of arrays: local, or a neighbour's halo column
  auto putv = [&](double *__restrict__ U0, double *__restrict__ V0, double *__restrict__ W0, double *__restrict__ UM,
                  double *__restrict__ VM, double *__restrict__ WM, long long q) {
    U0[q] = a; V0[q] = b; W0[q] = d;
    if (STEP3) { UM[q] = a; VM[q] = b; WM[q] = d; }
    if (top) {
      const long long qt = q + g.pk;
      U0[qt] = at; V0[qt] = bt; W0[qt] = 0.;
      if (STEP3) { UM[qt] = at; VM[qt] = bt; WM[qt] = 0.; }
    }
  };
  auto put = [&](long long q) {
    putv(u0, v0, w0, um, vm, wm, q);
    pres0[q] = pres0[q] + pc0;
  };
  put(c);
  const int ix = img_x(g, i), jy = img_y(g, j);
  if (ix >= 0) put(offF(g, ix, j, k));
  if (jy >= 0) { put(offF(g, i, jy, k)); if (ix >= 0) put(offF(g, ix, jy, k)); }
  if (XS >= 1) {
    // x is split over GPUs: the halo columns of p came from the neighbours (exchange before this kernel) and
    // pres0 += p has to reach the halo columns of pres0 too (src/modpois.f90:1096-1102)
    auto ph = [&](int hi) {
      const long long q = offF(g, hi, j, k);
      const double pv = p[q];
      pres0[q] = pres0[q] + pv;
      if (jy >= 0) { const long long q2 = offF(g, hi, jy, k); pres0[q2] = pres0[q2] + pv; }
    };
    if (i == 1) ph(0);
    if (i == g.imax) ph(g.imax + 1);
  }
}
// pres0 += p on the halo shell only (everything that is not an interior cell); the interior is done above.
__global__ void k_pres_update_shell(Geo g, const double *__restrict__ p, double *__restrict__ pres0) {
  const int lev = blockIdx.y;  // storage level 0 .. ktot+1
  const long long base = (long long)lev * g.pk;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (lev == 0 || lev == g.ktot + 1) {
    for (long long q = tid; q < g.pk; q += (long long)gridDim.x * blockDim.x) pres0[base + q] += p[base + q];
  } else {
    // ring: rows 0 and pj-1 (pi each), columns 0 and pi-1 of rows 1..pj-2
    const int nring = 2 * g.pi + 2 * (g.pj - 2);
    for (int r = tid; r < nring; r += gridDim.x * blockDim.x) {
      long long q;
      if (r < g.pi) q = r;
      else if (r < 2 * g.pi) q = (long long)(g.pj - 1) * g.pi + (r - g.pi);
      else { const int rr = r - 2 * g.pi; q = (long long)(1 + (rr >> 1)) * g.pi + ((rr & 1) ? g.pi - 1 : 0); }
      pres0[base + q] += p[base + q];
    }
  }
}

// tstep_integrate: src/modtstep.f90:171-340 for u,v,w.  ZERO: also zero the tendencies (:322-324);
// STEP3: um = u0 on the interior (halos are refreshed by the following halos call, :330-338).
template <bool ZERO, bool STEP3>
__global__ void __launch_bounds__(256) k_integrate(Geo g, double rk3coef, double *__restrict__ u0, double *__restrict__ v0,
                                                   double *__restrict__ w0, double *__restrict__ um, double *__restrict__ vm,
                                                   double *__restrict__ wm, double *__restrict__ up, double *__restrict__ vp,
                                                   double *__restrict__ wp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  if (i > g.imax || j > g.jmax) return;
  const long long c = offF(g, i, j, k), t = offT(g, i, j, k);
  const double a = um[c] + rk3coef * up[t];
  const double b = vm[c] + rk3coef * vp[t];
  const double d = wm[c] + rk3coef * wp[t];
  u0[c] = a; v0[c] = b; w0[c] = d;
  if (STEP3) { um[c] = a; vm[c] = b; wm[c] = d; }
  if (ZERO) { up[t] = 0.; vp[t] = 0.; wp[t] = 0.; }
}

// boundary (periodic x/y subset): src/modboundary.f90:163-204.  One thread per (si,sj) incl. halos.
__global__ void k_boundary_topbot(Geo g, double *__restrict__ u0, double *__restrict__ v0, double *__restrict__ w0,
                                  double *__restrict__ um, double *__restrict__ vm, double *__restrict__ wm) {
  const int si = blockIdx.x * blockDim.x + threadIdx.x;
  const int sj = blockIdx.y;
  if (si >= g.pi) return;
  const long long b = (long long)si + (long long)g.pi * sj;
  const long long k1 = b + g.pk * g.kh, kK = b + g.pk * (g.ktot + g.kh - 1), kT = kK + g.pk;
  wm[k1] = 0.; w0[k1] = 0.;
  if (g.BCtopm == 2) {
    um[kT] = 2 * g.Uinf - um[kK]; u0[kT] = 2 * g.Uinf - u0[kK];
    vm[kT] = 2 * g.Vinf - vm[kK]; v0[kT] = 2 * g.Vinf - v0[kK];
  } else {
    um[kT] = um[kK]; u0[kT] = u0[kK];
    vm[kT] = vm[kK]; v0[kT] = v0[kK];
  }
  w0[kT] = 0.; wm[kT] = 0.;
}
// reassure_fluxtop_boundary (src/modboundary.f90:392-431) for freeslip: u,v top ghost = top level
__global__ void k_fluxtop_uv(Geo g, double *__restrict__ u0, double *__restrict__ v0, double *__restrict__ um,
                             double *__restrict__ vm) {
  const int si = blockIdx.x * blockDim.x + threadIdx.x;
  const int sj = blockIdx.y;
  if (si >= g.pi) return;
  const long long b = (long long)si + (long long)g.pi * sj;
  const long long kK = b + g.pk * (g.ktot + g.kh - 1), kT = kK + g.pk;
  um[kT] = um[kK]; u0[kT] = u0[kK]; vm[kT] = vm[kK]; v0[kT] = v0[kK];
}

// ---------------------------------------------------------------------------------------------
// reductions.  Block-level max/sum in shared memory, then one atomic per block on doubles
// encoded so that atomicMax on the bit pattern works (all candidates are >= 0).
__device__ __forceinline__ double warp_max(double v) {
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v) {
  atomicMax((unsigned long long *)addr, (unsigned long long)__double_as_longlong(v));
}

// tstep_update maxima: src/modtstep.f90:113-127.  out[0] = courant, out[1] = diffusion number
// (both divided by dt later on the host side: they are linear in dt).
constexpr int CFL_KC = 16;   // levels per thread: 16 times fewer block reductions and atomics than one cell per thread
__global__ void __launch_bounds__(256) k_cfl(Geo g, const double *__restrict__ um, const double *__restrict__ vm,
                                             const double *__restrict__ wm, const double *__restrict__ ekm,
                                             const double *__restrict__ ekh, double dt, double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k0 = blockIdx.z * CFL_KC + 1, k1 = min(k0 + CFL_KC, g.ktot + 1);
  double c = 0., d = 0.;
  if (i <= g.imax && j <= g.jmax) {
    long long q = offF(g, i, j, k0);
    for (int k = k0; k < k1; k++, q += g.pk) {       // maxima: any order gives the same bits
      c = fmax(c, (fabs(um[q]) * g.dxi + fabs(vm[q]) * g.dyi + fabs(wm[q]) / g.dzh[k]) * dt);
      const double m = (g.dzh2i[k] + g.dx2i + g.dy2i) * dt;
      d = fmax(d, fmax(ekm[q] * m, ekh[q] * m));
    }
  }
  c = warp_max(c); d = warp_max(d);
  __shared__ double sc[8], sd[8];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if ((tid & 31) == 0) { sc[tid >> 5] = c; sd[tid >> 5] = d; }
  __syncthreads();
  if (tid < 32) {
    const int nw = (blockDim.x * blockDim.y + 31) >> 5;
    c = tid < nw ? sc[tid] : 0.; d = tid < nw ? sd[tid] : 0.;
    c = warp_max(c); d = warp_max(d);
    if (tid == 0) { atomic_max_nonneg(out, c); atomic_max_nonneg(out + 1, d); }
  }
}

// chkdiv: src/modchecksim.f90:161-203.  out[0] = max|div|, out[1] = sum div*dx*dy*dzf, out[2] = sum div^2
__global__ void __launch_bounds__(256) k_div(Geo g, const double *__restrict__ u0, const double *__restrict__ v0,
                                             const double *__restrict__ w0, double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  double m = 0., s = 0., s2 = 0.;
  if (i <= g.imax && j <= g.jmax) {
    const long long q = offF(g, i, j, k);
    const double div = (u0[q + 1] - u0[q]) * g.dxi + (v0[q + g.pi] - v0[q]) * g.dyi + (w0[q + g.pk] - w0[q]) * g.dzfi[k];
    m = fabs(div); s = div * g.dx * g.dy * g.dzf[k]; s2 = div * div;
  }
  m = warp_max(m); s = warp_sum(s); s2 = warp_sum(s2);
  __shared__ double sm[8], ss[8], ss2[8];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if ((tid & 31) == 0) { sm[tid >> 5] = m; ss[tid >> 5] = s; ss2[tid >> 5] = s2; }
  __syncthreads();
  if (tid < 32) {
    const int nw = (blockDim.x * blockDim.y + 31) >> 5;
    m = tid < nw ? sm[tid] : 0.; s = tid < nw ? ss[tid] : 0.; s2 = tid < nw ? ss2[tid] : 0.;
    m = warp_max(m); s = warp_sum(s); s2 = warp_sum(s2);
    if (tid == 0) { atomic_max_nonneg(out, m); atomicAdd(out + 1, s); atomicAdd(out + 2, s2); }
  }
}

}  // namespace udg
