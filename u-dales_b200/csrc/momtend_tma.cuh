// momtend_tma.cuh — fused momentum advection + diffusion (advecu/v/w_2nd + diffu/v/w,
// src/modadvection.f90:158-314, src/modsubgrid.f90:672-997) as ONE persistent sm_100a kernel.
//
// Design:
//  * flux form.  Every term of the three tendencies is a difference of a face/edge quantity
//      Gxx, Gyy, Gzz   (cell centres)    = 2 K dUi/dxi - 1/4 (Ui + Ui')^2
//      Cxy, Cxz, Cyz   (cell edges)      = K_edge (dUi/dxj + dUj/dxi) - 1/4 (Ui + Ui')(Uj + Uj')
//    and each quantity is shared by two equations and two cells (e.g. the z-flux of u through the
//    edge (i-1/2,k-1/2) is the x-flux of w through the same edge).  Computing each once and
//    exchanging it through shared memory cuts the arithmetic from ~295 to ~100 flop/cell, which
//    moves the kernel off the fp64 ridge onto the HBM roof.
//  * k-marching.  A CTA owns a 32x16 (i,j) tile and walks up in k.  Planes of u0,v0,w0,pres0,ekm
//    (tile + 1 halo) arrive by TMA (cp.async.bulk.tensor.3d) into a ring of shared-memory stages,
//    completion through mbarriers; values of level k+1 loaded at step k are carried in registers
//    to serve as level k at step k+1, vertical-edge fluxes likewise.
//  * 512 "main" threads own one cell column each; 2 "ring" warps compute the edge quantities of the
//    left column / bottom row just outside the tile so tiles need no overlap and stores stay aligned.
//  * persistent: each CTA walks a static list of (tile, k-chunk) items as one flat plane stream, so
//    the TMA ring never drains between items.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace udg {

constexpr int MT_TX = 32, MT_TY = 16, MT_S = 4;
constexpr int MT_BX = MT_TX + 2, MT_BY = MT_TY + 2;
constexpr int MT_BOX_BYTES = MT_BX * MT_BY * 8;
constexpr int MT_BOX_PAD = (MT_BOX_BYTES + 127) / 128 * 128;
constexpr int MT_NF = 5;  // u v w p e
constexpr int MT_STAGE_BYTES = MT_NF * MT_BOX_PAD;
constexpr int MT_XP = MT_BX;                 // pitch of the exchange arrays
constexpr int MT_XARR = MT_XP * (MT_TY + 1);  // elements per exchange array
constexpr int MT_NX = 5;                     // Gxx Gyy Cxy CxzTop CyzTop
constexpr int MT_MAIN = MT_TX * MT_TY, MT_RING = 64, MT_THREADS = MT_MAIN + MT_RING;
constexpr int MT_SMEM = MT_S * MT_STAGE_BYTES + 2 * MT_NX * MT_XARR * 8 + 64;

struct MomTmaParams {
  Geo g;
  int ntx, nty, nchunk;   // tiles in i, j; chunks in k
  int nitems;
  double *up, *vp, *wp;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

// item -> tile / level range.  Levels are split as evenly as possible over nchunk chunks.
__device__ __forceinline__ void mt_item(const MomTmaParams &P, int item, int &it, int &jt, int &k0, int &k1) {
  const int ntile = P.ntx * P.nty;
  const int c = item / ntile, t = item - c * ntile;
  jt = t / P.ntx;
  it = t - jt * P.ntx;
  const int K = P.g.ktot;
  k0 = 1 + (int)(((long long)K * c) / P.nchunk);
  k1 = 1 + (int)(((long long)K * (c + 1)) / P.nchunk);  // exclusive
}

template <bool ADV, bool DIFF, bool LES, bool ACC>
__global__ void __launch_bounds__(MT_THREADS, 1)
    k_momtend_tma(const __grid_constant__ CUtensorMap mu, const __grid_constant__ CUtensorMap mv,
                  const __grid_constant__ CUtensorMap mw, const __grid_constant__ CUtensorMap mp,
                  const __grid_constant__ CUtensorMap me, const MomTmaParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  double *xch = reinterpret_cast<double *>(smem + MT_S * MT_STAGE_BYTES);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + MT_S * MT_STAGE_BYTES + 2 * MT_NX * MT_XARR * 8);
  const Geo &g = P.g;
  const int tid = threadIdx.x;

  // ---- role ----
  int gx, gy;
  bool gpoint = true;
  const bool is_main = tid < MT_MAIN;
  if (is_main) { gx = tid & (MT_TX - 1); gy = tid / MT_TX; }
  else {
    const int r = tid - MT_MAIN;
    if (r < MT_TX) { gx = r; gy = -1; }
    else if (r < MT_TX + MT_TY) { gx = -1; gy = r - MT_TX; }
    else if (r == MT_TX + MT_TY) { gx = -1; gy = -1; }
    else { gx = 0; gy = 0; gpoint = false; }
  }
  const int idx0 = (gy + 1) * MT_BX + (gx + 1);   // own cell in a plane box
  const int xidx = (gy + 1) * MT_XP + (gx + 1);   // own slot in an exchange array

  if (tid == 0) {
    for (int s = 0; s < MT_S; s++) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // ---- producer state (thread 0 only): flat plane stream over this CTA's items ----
  int p_item = blockIdx.x, p_pl = 0, p_np = 0, p_it = 0, p_jt = 0, p_k0 = 0;
  int issued = 0, released = -1;
  if (tid == 0 && p_item < P.nitems) {
    int k1;
    mt_item(P, p_item, p_it, p_jt, p_k0, k1);
    p_np = k1 - p_k0 + 2;
  }
  auto issue_ready = [&]() {
    // issue every plane whose stage has been released by the consumers
    while (p_item < P.nitems && issued - MT_S <= released) {
      const int st = issued % MT_S;
      unsigned char *dst = smem + st * MT_STAGE_BYTES;
      uint64_t *bar = &bars[st];
      mbar_expect_tx(bar, MT_NF * MT_BOX_BYTES);
      const int c0 = p_it * MT_TX, c1 = p_jt * MT_TY, c2 = p_k0 - 1 + p_pl;
      tma_load_3d(dst + 0 * MT_BOX_PAD, &mu, c0, c1, c2, bar);
      tma_load_3d(dst + 1 * MT_BOX_PAD, &mv, c0, c1, c2, bar);
      tma_load_3d(dst + 2 * MT_BOX_PAD, &mw, c0, c1, c2, bar);
      tma_load_3d(dst + 3 * MT_BOX_PAD, &mp, c0, c1, c2, bar);
      tma_load_3d(dst + 4 * MT_BOX_PAD, &me, c0, c1, c2, bar);
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

  const double dxi = g.dxi, dyi = g.dyi;
  const double numol = g.numol;
  int q = 0;  // stream index of the current item's first plane
  int gstep = 0;  // running step counter: parity selects the exchange buffer (never reused back-to-back)
  for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
    int it, jt, k0, k1;
    mt_item(P, item, it, jt, k0, k1);
    const int np = k1 - k0 + 2;
    const int ci = it * MT_TX + gx + 1, cj = jt * MT_TY + gy + 1;  // Fortran cell indices of this thread
    const bool store_ok = is_main && ci <= g.imax && cj <= g.jmax;
    const long long tbase = (long long)ci + (long long)g.pi * cj;  // storage (ih=jh=1): col ci, row cj

    // carried registers
    double u_ip = 0, v_jp = 0, w_c = 0, e_c = numol, e_ip = numol, e_jp = numol, p_c = 0, p_km = 0;
    double cxz_bot = 0, cyz_bot = 0, cxz_botL = 0, cyz_botB = 0, gzz_prev = 0;

    mbar_wait(&bars[q % MT_S], (q / MT_S) & 1);
    {
      const double *b = reinterpret_cast<const double *>(smem + (q % MT_S) * MT_STAGE_BYTES);
      if (gpoint) {
        u_ip = b[0 * (MT_BOX_PAD / 8) + idx0 + 1];
        v_jp = b[1 * (MT_BOX_PAD / 8) + idx0 + MT_BX];
        w_c = b[2 * (MT_BOX_PAD / 8) + idx0];
        if (ADV) p_c = b[3 * (MT_BOX_PAD / 8) + idx0];
        if (DIFF && LES) {
          e_c = b[4 * (MT_BOX_PAD / 8) + idx0];
          e_ip = b[4 * (MT_BOX_PAD / 8) + idx0 + 1];
          e_jp = b[4 * (MT_BOX_PAD / 8) + idx0 + MT_BX];
        }
      }
    }

    for (int s = 0; s < np - 1; s++) {
      const int k = k0 - 1 + s, K = k + 1;
      const int qc = q + s, qn = q + s + 1;
      mbar_wait(&bars[qn % MT_S], (qn / MT_S) & 1);
      const double *bc = reinterpret_cast<const double *>(smem + (qc % MT_S) * MT_STAGE_BYTES);
      const double *bn = reinterpret_cast<const double *>(smem + (qn % MT_S) * MT_STAGE_BYTES);
      double *X = xch + (gstep & 1) * (MT_NX * MT_XARR);
      gstep++;

      const double dzfk = __ldg(g.dzf + k), dzfK = __ldg(g.dzf + K);
      const double dzhiK = __ldg(g.dzhi + K), dzfik = __ldg(g.dzfi + k);

      double Gxx = 0, Gyy = 0, Cxy = 0, CxzT = 0, CyzT = 0, Gzz = 0;
      double uK_ip = 0, vK_jp = 0, wK = 0, eK = numol, eK_ip = numol, eK_jp = numol, pK = 0, p_im = 0, p_jm = 0;
      if (gpoint) {
        // ---- phase A: loads ----
        uK_ip = bn[0 * (MT_BOX_PAD / 8) + idx0 + 1];
        vK_jp = bn[1 * (MT_BOX_PAD / 8) + idx0 + MT_BX];
        wK = bn[2 * (MT_BOX_PAD / 8) + idx0];
        const double wK_ip = bn[2 * (MT_BOX_PAD / 8) + idx0 + 1];
        const double wK_jp = bn[2 * (MT_BOX_PAD / 8) + idx0 + MT_BX];
        const double u_c = bc[0 * (MT_BOX_PAD / 8) + idx0];
        const double v_c = bc[1 * (MT_BOX_PAD / 8) + idx0];
        const double u_ipjp = bc[0 * (MT_BOX_PAD / 8) + idx0 + MT_BX + 1];
        const double v_ipjp = bc[1 * (MT_BOX_PAD / 8) + idx0 + MT_BX + 1];
        double e_ipjp = numol;
        if (DIFF && LES) {
          eK = bn[4 * (MT_BOX_PAD / 8) + idx0];
          eK_ip = bn[4 * (MT_BOX_PAD / 8) + idx0 + 1];
          eK_jp = bn[4 * (MT_BOX_PAD / 8) + idx0 + MT_BX];
          e_ipjp = bc[4 * (MT_BOX_PAD / 8) + idx0 + MT_BX + 1];
        }
        if (ADV && is_main) {
          pK = bn[3 * (MT_BOX_PAD / 8) + idx0];
          p_im = bc[3 * (MT_BOX_PAD / 8) + idx0 - 1];
          p_jm = bc[3 * (MT_BOX_PAD / 8) + idx0 - MT_BX];
        }
        // ---- phase A: face / edge quantities ----
        const double su = u_c + u_ip, sv = v_c + v_jp, sw = w_c + wK;
        const double sxy_u = u_ipjp + u_ip, sxy_v = v_ipjp + v_jp;
        const double axz_u = (uK_ip * dzfk + u_ip * dzfK) * dzhiK, axz_w = wK_ip + wK;
        const double ayz_v = (vK_jp * dzfk + v_jp * dzfK) * dzhiK, ayz_w = wK_jp + wK;
        if (ADV) {
          Gxx = -0.25 * su * su;
          Gyy = -0.25 * sv * sv;
          Gzz = -0.25 * sw * sw;
          Cxy = -0.25 * sxy_u * sxy_v;
          CxzT = -0.25 * axz_u * axz_w;
          CyzT = -0.25 * ayz_v * ayz_w;
        }
        if (DIFF) {
          const double dzhiqK = 0.25 * dzhiK;
          double e4, exK, eyK;
          if (LES) {
            e4 = 0.25 * ((e_c + e_ip) + (e_jp + e_ipjp));
            exK = (dzfk * (eK_ip + eK) + dzfK * (e_ip + e_c)) * dzhiqK;
            eyK = (dzfk * (eK_jp + eK) + dzfK * (e_jp + e_c)) * dzhiqK;
          } else {
            e4 = exK = eyK = numol;
          }
          Gxx += 2. * dxi * e_c * (u_ip - u_c);
          Gyy += 2. * dyi * e_c * (v_jp - v_c);
          Gzz += 2. * dzfik * e_c * (wK - w_c);
          Cxy += e4 * ((u_ipjp - u_ip) * dyi + (v_ipjp - v_jp) * dxi);
          CxzT += exK * ((uK_ip - u_ip) * dzhiK + (wK_ip - wK) * dxi);
          CyzT += eyK * ((vK_jp - v_jp) * dzhiK + (wK_jp - wK) * dyi);
        }
        X[0 * MT_XARR + xidx] = Gxx;
        X[1 * MT_XARR + xidx] = Gyy;
        X[2 * MT_XARR + xidx] = Cxy;
        X[3 * MT_XARR + xidx] = CxzT;
        X[4 * MT_XARR + xidx] = CyzT;
      }
      __syncthreads();
      if (tid == 0) {
        // plane qc has been fully read (all its loads precede the barrier); at the last step of the item
        // plane qn is done as well (its values live on in registers only until the item ends)
        released = (s == np - 2) ? qn : qc;
        issue_ready();
      }
      if (is_main) {
        // ---- phase B: assemble the three tendencies ----
        const double GxxL = X[0 * MT_XARR + xidx - 1];
        const double GyyB = X[1 * MT_XARR + xidx - MT_XP];
        const double CxyL = X[2 * MT_XARR + xidx - 1];
        const double CxyB = X[2 * MT_XARR + xidx - MT_XP];
        const double CxyBL = X[2 * MT_XARR + xidx - MT_XP - 1];
        const double CxzTL = X[3 * MT_XARR + xidx - 1];
        const double CyzTB = X[4 * MT_XARR + xidx - MT_XP];
        if (s >= 1 && store_ok) {
          const long long t = tbase + (long long)g.pk * (k - 1);
          double ru = (Gxx - GxxL) * dxi + (CxyL - CxyBL) * dyi + (CxzTL - cxz_botL) * dzfik;
          double rv = (CxyB - CxyBL) * dxi + (Gyy - GyyB) * dyi + (CyzTB - cyz_botB) * dzfik;
          if (ADV) {
            ru -= (p_c - p_im) * dxi;
            rv -= (p_c - p_jm) * dyi;
          }
          if (ACC) { ru += P.up[t]; rv += P.vp[t]; }
          P.up[t] = ru;
          P.vp[t] = rv;
          if (k >= 2) {
            const double dzhik = __ldg(g.dzhi + k);
            double rw = (cxz_bot - cxz_botL) * dxi + (cyz_bot - cyz_botB) * dyi + (Gzz - gzz_prev) * dzhik;
            if (ADV) rw -= (p_c - p_km) * dzhik;
            if (ACC) rw += P.wp[t];
            P.wp[t] = rw;
          } else if (!ACC) {
            P.wp[t] = 0.0;
          }
        }
        cxz_botL = CxzTL;
        cyz_botB = CyzTB;
      }
      // ---- carry ----
      cxz_bot = CxzT; cyz_bot = CyzT; gzz_prev = Gzz;
      p_km = p_c; p_c = pK;
      u_ip = uK_ip; v_jp = vK_jp; w_c = wK;
      e_c = eK; e_ip = eK_ip; e_jp = eK_jp;
    }
    q += np;
  }
}

}  // namespace udg
