// closure_tma.cuh — Vreman (2004) eddy viscosity (closure, src/modsubgrid.f90:269-360) as a persistent
// TMA-staged k-marching kernel: same tile / plane-stream machinery as k_momtend_tma, three input fields
// (u0,v0,w0: 24 B/cell) and two outputs (ekm, ekh: 16 B/cell).  Each thread owns one (i,j) column;
// level k+1 operands loaded at step k are carried in registers as level k (and their vertical sums as
// level k-1), so every plane box is read from shared memory once.  Operand order inside each gradient
// follows the reference expression, so results differ from it by FMA contraction only — important
// next to the hard `bb < 1e-8` switch (:323).
#pragma once
#include "momtend_tma.cuh"
#include "stencil_v1.cuh"

namespace udg {

constexpr int CL_NF = 3;
constexpr int CL_S = 4;
constexpr int CL_STAGE_BYTES = CL_NF * MT_BOX_PAD;
constexpr int CL_THREADS = MT_TX * MT_TY;
constexpr int CL_SMEM = CL_S * CL_STAGE_BYTES + 64;

template <int MINB>
__global__ void __launch_bounds__(CL_THREADS, MINB)
    k_closure_vreman_tma(const __grid_constant__ CUtensorMap mu, const __grid_constant__ CUtensorMap mv,
                         const __grid_constant__ CUtensorMap mw, const MomTmaParams P, double *__restrict__ ekm,
                         double *__restrict__ ekh, int halo, PeerCols pc) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + CL_S * CL_STAGE_BYTES);
  const Geo &g = P.g;
  const int tid = threadIdx.x;
  const int gx = tid & (MT_TX - 1), gy = tid / MT_TX;
  const int idx0 = (gy + 1) * MT_BX + (gx + 1);
  constexpr int FU = 0, FV = MT_BOX_PAD / 8, FW = 2 * (MT_BOX_PAD / 8);

  if (tid == 0) {
    for (int s = 0; s < CL_S; s++) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  int p_item = blockIdx.x, p_pl = 0, p_np = 0, p_it = 0, p_jt = 0, p_k0 = 0;
  int issued = 0, released = -1;
  if (tid == 0 && p_item < P.nitems) {
    int k1;
    mt_item(P, p_item, p_it, p_jt, p_k0, k1);
    p_np = k1 - p_k0 + 2;
  }
  auto issue_ready = [&]() {
    while (p_item < P.nitems && issued - CL_S <= released) {
      const int st = issued % CL_S;
      unsigned char *dst = smem + st * CL_STAGE_BYTES;
      uint64_t *bar = &bars[st];
      mbar_expect_tx(bar, CL_NF * MT_BOX_BYTES);
      const int c0 = p_it * MT_TX, c1 = p_jt * MT_TY, c2 = p_k0 - 1 + p_pl;
      tma_load_3d(dst + 0 * MT_BOX_PAD, &mu, c0, c1, c2, bar);
      tma_load_3d(dst + 1 * MT_BOX_PAD, &mv, c0, c1, c2, bar);
      tma_load_3d(dst + 2 * MT_BOX_PAD, &mw, c0, c1, c2, bar);
      issued++;
      if (++p_pl == p_np) {
        p_item += gridDim.x;
        p_pl = 0;
        if (p_item < P.nitems) {
          int k1;
          mt_item(P, p_item, p_it, p_jt, p_k0, k1);
          p_np = k1 - p_k0 + 2;
        }
      }
    }
  };
  if (tid == 0) issue_ready();

  const double dxi = g.dxi, dyi = g.dyi, dxiq = g.dxiq, dyiq = g.dyiq, dx2 = g.dx2, dy2 = g.dy2;
  int q = 0;
  for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
    int it, jt, k0, k1;
    mt_item(P, item, it, jt, k0, k1);
    const int np = k1 - k0 + 2;
    const int ci = it * MT_TX + gx + 1, cj = jt * MT_TY + gy + 1;
    const bool store_ok = ci <= g.imax && cj <= g.jmax;

    mbar_wait(&bars[q % CL_S], (q / CL_S) & 1);
    const double *b0 = reinterpret_cast<const double *>(smem + (q % CL_S) * CL_STAGE_BYTES);
    // level k slots (initially level k0-1)
    double u_c = b0[FU + idx0], u_ip = b0[FU + idx0 + 1];
    double v_c = b0[FV + idx0], v_jp = b0[FV + idx0 + MT_BX];
    double w_c = b0[FW + idx0], w_ip = b0[FW + idx0 + 1], w_im = b0[FW + idx0 - 1];
    double w_jp = b0[FW + idx0 + MT_BX], w_jm = b0[FW + idx0 - MT_BX];
    double su_km = 0, sv_km = 0;

    for (int s = 0; s < np - 1; s++) {
      const int k = k0 - 1 + s, K = k + 1;
      const int qc = q + s, qn = q + s + 1;
      mbar_wait(&bars[qn % CL_S], (qn / CL_S) & 1);
      const double *bc = reinterpret_cast<const double *>(smem + (qc % CL_S) * CL_STAGE_BYTES);
      const double *bn = reinterpret_cast<const double *>(smem + (qn % CL_S) * CL_STAGE_BYTES);
      const double uK_c = bn[FU + idx0], uK_ip = bn[FU + idx0 + 1];
      const double vK_c = bn[FV + idx0], vK_jp = bn[FV + idx0 + MT_BX];
      const double wK_c = bn[FW + idx0], wK_ip = bn[FW + idx0 + 1], wK_im = bn[FW + idx0 - 1];
      const double wK_jp = bn[FW + idx0 + MT_BX], wK_jm = bn[FW + idx0 - MT_BX];
      const double su_k = u_ip + u_c, sv_k = v_jp + v_c;
      if (s >= 1) {
        const double u_ipjp = bc[FU + idx0 + MT_BX + 1], u_jp = bc[FU + idx0 + MT_BX];
        const double u_ipjm = bc[FU + idx0 - MT_BX + 1], u_jm = bc[FU + idx0 - MT_BX];
        const double v_ipjp = bc[FV + idx0 + MT_BX + 1], v_ip = bc[FV + idx0 + 1];
        const double v_imjp = bc[FV + idx0 + MT_BX - 1], v_im = bc[FV + idx0 - 1];
        const double dzfk = __ldg(g.dzf + k), dzfK = __ldg(g.dzf + K), dzfkm = __ldg(g.dzf + k - 1);
        const double dzhik = __ldg(g.dzhi + k), dzhiK = __ldg(g.dzhi + K);
        const double dzfiqk = __ldg(g.dzfiq + k), dzfik = __ldg(g.dzfi + k), dzf2 = __ldg(g.dzf2 + k);
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
        if (store_ok) ek_store<false>(g, ci, cj, k, e, ekm, ekh, halo, pc);
      }
      __syncthreads();
      if (tid == 0) {
        released = (s == np - 2) ? qn : qc;
        issue_ready();
      }
      su_km = su_k; sv_km = sv_k;
      u_c = uK_c; u_ip = uK_ip; v_c = vK_c; v_jp = vK_jp;
      w_c = wK_c; w_ip = wK_ip; w_im = wK_im; w_jp = wK_jp; w_jm = wK_jm;
    }
    q += np;
  }
}

}  // namespace udg
