// poisson_xz.cuh — the x-direction part of the Poisson solve as ONE pass over memory:
//   forward real FFT in x  ->  tridiagonal solve in z  ->  inverse real FFT in x
// (src/modpois.f90:478-490 + solmpj :1107-1166 + :669-679) for power-of-two itot.
//
// One CTA owns one x-z plane (fixed j: one packed y-slot after the forward y transform) and walks it in
// batches of LANES levels.  A batch (LANES contiguous x-lines) is staged into shared memory, transformed by
// the same four-step register FFT as k_rfft_fast, and then every thread runs the Thomas recurrence of "its"
// packed x-slot(s) through the LANES levels of the batch, carrying x'_{k-1} in a register from batch to batch.
// The forward-swept batch goes back to the work array in place; the downward pass re-reads it (those lines
// were written last and are the most likely L2 residents), runs the back substitution, the inverse transform
// and writes the solution.  The top batch never leaves shared memory.  Compared to the three separate passes
// (x-FFT, z-solve with its forward + backward sweeps over HBM, inverse x-FFT: 16 + 32 + 16 B/cell of traffic)
// this touches HBM for 16 B/cell plus whatever part of the x' round trip misses L2.
//
// Factors 1/(b_k + lambda - a_k d_{k-1}) come from the table of k_zfactor (poisson_fast.cuh).
#pragma once
#include "poisson_fast.cuh"

namespace udg {

template <int R1, int R2, int LANES>
struct XzT {
  static constexpr int H = R1 * R2, N = 2 * H, NT = LANES * R2, NK1 = R1 / R2, NPAIR = (H / 2) / R2 + 1;
  static constexpr int NIT = (LANES * H) / NT;        // staging iterations per thread (= R1)
  static constexpr int NSLOT = (N + NT - 1) / NT;     // packed x-slots per thread in the z recurrence
  static constexpr int SMEM = LANES * (H + 1) * 16;
};

#define XSA(p) (lane * (H + 1) + (p))
#define XRS(r) (lane * 2 * (H + 1) + (r))

// passes 1 + 2 of the four-step complex FFT of length H = R1 R2 (one line per lane, R2 threads per line)
template <int R1, int R2, int LANES, bool INV>
__device__ __forceinline__ void xz_fourstep(double2 (&v)[R1], double2 (&u)[R1 / R2][R2], double2 *buf, const double2 *__restrict__ tw,
                                            int lane, int j) {
  constexpr int H = R1 * R2, NK1 = R1 / R2;
  dft_reg<R1, INV>(v);
#pragma unroll
  for (int k1 = 0; k1 < R1; k1++) {
    double2 y = v[brev<R1>(k1)];
    if (k1 != 0) {
      double2 w = tw[2 * j * k1];
      if (INV) w.y = -w.y;
      y = cmul(y, w);
    }
    buf[XSA(k1 * R2 + j)] = y;
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < NK1; t++) {
    const int k1 = j + R2 * t;
#pragma unroll
    for (int jj = 0; jj < R2; jj++) u[t][jj] = buf[XSA(k1 * R2 + jj)];
    dft_reg<R2, INV>(u[t]);
  }
}

// natural real lines (as double2 pairs) in the tile -> packed half-complex reals [Re0,Re1,Im1,...,Re(n/2)] / sqrt(n), in place
template <int R1, int R2, int LANES>
__device__ __forceinline__ void xz_tile_fwd(double2 *buf, const double2 *__restrict__ tw, double fac, int lane, int j) {
  using T = XzT<R1, R2, LANES>;
  constexpr int H = T::H, N = T::N, NK1 = T::NK1, NPAIR = T::NPAIR;
  double *rbuf = reinterpret_cast<double *>(buf);
  double2 v[R1], u[NK1][R2];
#pragma unroll
  for (int q = 0; q < R1; q++) v[q] = buf[XSA(j + R2 * q)];
  __syncthreads();
  xz_fourstep<R1, R2, LANES, false>(v, u, buf, tw, lane, j);
  __syncthreads();
#pragma unroll
  for (int t = 0; t < NK1; t++)
#pragma unroll
    for (int k2 = 0; k2 < R2; k2++) buf[XSA(j + R2 * t + R1 * k2)] = u[t][brev<R2>(k2)];
  __syncthreads();
  double2 xk[NPAIR], xh[NPAIR];
#pragma unroll
  for (int t = 0; t < NPAIR; t++) {
    const int k = j + R2 * t;
    xk[t] = xh[t] = make_double2(0., 0.);
    if (k <= H / 2) {
      const double2 Zk = buf[XSA(k)];
      if (k == 0) {
        xk[t] = make_double2((Zk.x + Zk.y) * fac, (Zk.x - Zk.y) * fac);
      } else {
        const double2 Zc = cconj(buf[XSA(H - k)]);
        const double2 E = make_double2(0.5 * (Zk.x + Zc.x), 0.5 * (Zk.y + Zc.y));
        const double2 D = csub(Zk, Zc);
        const double2 w = tw[k];
        const double2 Tt = cmul(make_double2(0.5 * w.y, -0.5 * w.x), D);
        const double2 a = cadd(E, Tt), b = cconj(csub(E, Tt));
        xk[t] = make_double2(a.x * fac, a.y * fac);
        xh[t] = make_double2(b.x * fac, b.y * fac);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < NPAIR; t++) {
    const int k = j + R2 * t;
    if (k <= H / 2) {
      if (k == 0) { rbuf[XRS(0)] = xk[t].x; rbuf[XRS(N - 1)] = xk[t].y; }
      else {
        rbuf[XRS(2 * k - 1)] = xk[t].x; rbuf[XRS(2 * k)] = xk[t].y;
        rbuf[XRS(2 * (H - k) - 1)] = xh[t].x; rbuf[XRS(2 * (H - k))] = xh[t].y;
      }
    }
  }
}

// packed half-complex reals in the tile -> natural real lines (as double2 pairs) / sqrt(n), in place
template <int R1, int R2, int LANES>
__device__ __forceinline__ void xz_tile_inv(double2 *buf, const double2 *__restrict__ tw, double fac, int lane, int j) {
  using T = XzT<R1, R2, LANES>;
  constexpr int H = T::H, N = T::N, NK1 = T::NK1, NPAIR = T::NPAIR;
  double *rbuf = reinterpret_cast<double *>(buf);
  double2 zk[NPAIR], zh[NPAIR];
#pragma unroll
  for (int t = 0; t < NPAIR; t++) {
    const int k = j + R2 * t;
    zk[t] = zh[t] = make_double2(0., 0.);
    if (k <= H / 2) {
      if (k == 0) {
        const double x0 = rbuf[XRS(0)], y0 = rbuf[XRS(N - 1)];
        zk[t] = make_double2(x0 + y0, x0 - y0);
      } else {
        const double x0 = rbuf[XRS(2 * k - 1)], x1 = rbuf[XRS(2 * k)];
        const double y0 = rbuf[XRS(2 * (H - k) - 1)], y1 = rbuf[XRS(2 * (H - k))];
        const double2 A = make_double2(x0 + y0, x1 - y1), Bv = make_double2(x0 - y0, x1 + y1);
        const double2 w = tw[k];
        const double2 Tt = cmul(make_double2(w.y, w.x), Bv);
        zk[t] = cadd(A, Tt);
        zh[t] = cconj(csub(A, Tt));
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < NPAIR; t++) {
    const int k = j + R2 * t;
    if (k <= H / 2) {
      if (k != 0 && k != H - k) buf[XSA(H - k)] = zh[t];
      buf[XSA(k)] = zk[t];
    }
  }
  __syncthreads();
  double2 v[R1], u[NK1][R2];
#pragma unroll
  for (int q = 0; q < R1; q++) v[q] = buf[XSA(j + R2 * q)];
  __syncthreads();
  xz_fourstep<R1, R2, LANES, true>(v, u, buf, tw, lane, j);
  __syncthreads();
#pragma unroll
  for (int t = 0; t < NK1; t++)
#pragma unroll
    for (int k2 = 0; k2 < R2; k2++) {
      const double2 z = u[t][brev<R2>(k2)];
      buf[XSA(j + R2 * t + R1 * k2)] = make_double2(z.x * fac, z.y * fac);
    }
}

// x: halo-free work array (n, nj, K) holding the y-transformed right-hand side; solved in place.
// xs_j / xs_k: element strides between planes / levels.  j0g: global packed y-slot of local plane 0.
template <int R1, int R2, int LANES, int MINB>
__global__ void __launch_bounds__(LANES *R2, MINB) k_xzsolve(const double2 *__restrict__ tw, double *x, long long xs_j, long long xs_k, int K,
                                                        int j0g, int nxh, int nyh, const double *__restrict__ zt,
                                                        const double *__restrict__ a, const double *__restrict__ c, double fac) {
  using T = XzT<R1, R2, LANES>;
  constexpr int H = T::H, N = T::N, NT = T::NT, NIT = T::NIT, NSLOT = T::NSLOT;
  constexpr int CH = LANES > 8 ? 8 : LANES;   // levels per register batch of the recurrence
  extern __shared__ double2 buf[];
  double *rbuf = reinterpret_cast<double *>(buf);
  const int lane = threadIdx.x, j = threadIdx.y, tid = j * LANES + lane;
  double *xp = x + (long long)blockIdx.x * xs_j;
  const int jy = (j0g + (int)blockIdx.x + 1) >> 1;
  const long long tk = (long long)nxh * nyh;
  const double *ztp = zt + (long long)jy * nxh;
  const bool al = ((((size_t)xp) & 15) == 0) && ((xs_k & 1) == 0);

  // LANES contiguous lines (levels k0 .. k0+nb-1) <-> the padded tile; all loads of a thread in flight at once.
  // The lines re-read on the way down were written by other threads of this CTA: ld.global.cg, never the
  // non-coherent path.
  constexpr int SCH = NIT > 8 ? 8 : NIT;   // requests in flight per thread and staging round
  auto stage_in = [&](int k0, int nb) {
#pragma unroll
    for (int i0 = 0; i0 < NIT; i0 += SCH) {
      double2 st[SCH];
#pragma unroll
      for (int it = 0; it < SCH; it++) {
        const int idx = tid + (i0 + it) * NT;
        if (idx < nb * H) {
          const int b = idx / H, m = idx - b * H;
          const double *q = xp + (long long)(k0 + b) * xs_k + 2 * m;
          st[it] = al ? __ldcg(reinterpret_cast<const double2 *>(q)) : make_double2(__ldcg(q), __ldcg(q + 1));
        }
      }
#pragma unroll
      for (int it = 0; it < SCH; it++) {
        const int idx = tid + (i0 + it) * NT;
        if (idx < nb * H) {
          const int b = idx / H, m = idx - b * H;
          buf[b * (H + 1) + m] = st[it];
        }
      }
    }
  };
  auto stage_out = [&](int k0, int nb) {
#pragma unroll
    for (int i0 = 0; i0 < NIT; i0 += SCH) {
      double2 st[SCH];
#pragma unroll
      for (int it = 0; it < SCH; it++) {
        const int idx = tid + (i0 + it) * NT;
        if (idx < nb * H) {
          const int b = idx / H, m = idx - b * H;
          st[it] = buf[b * (H + 1) + m];
        }
      }
#pragma unroll
      for (int it = 0; it < SCH; it++) {
        const int idx = tid + (i0 + it) * NT;
        if (idx < nb * H) {
          const int b = idx / H, m = idx - b * H;
          double *q = xp + (long long)(k0 + b) * xs_k + 2 * m;
          if (al) *reinterpret_cast<double2 *>(q) = st[it];
          else { q[0] = st[it].x; q[1] = st[it].y; }
        }
      }
    }
  };

  double carry[NSLOT];
#pragma unroll
  for (int si = 0; si < NSLOT; si++) carry[si] = 0.;
  const int nbat = (K + LANES - 1) / LANES;

  // ---------------- upward: forward transform + forward elimination (src/modpois.f90:1120-1155) ----------------
#pragma unroll 1
  for (int b = 0; b < nbat; b++) {
    const int k0 = b * LANES, nb = min(LANES, K - k0);
    stage_in(k0, nb);
    __syncthreads();
    xz_tile_fwd<R1, R2, LANES>(buf, tw, fac, lane, j);
    __syncthreads();
#pragma unroll
    for (int si = 0; si < NSLOT; si++) {
      const int s = tid + si * NT;
      if (s < N) {
        const double *zq = ztp + ((s + 1) >> 1) + (long long)k0 * tk;
        double xprev = carry[si];
#pragma unroll 1
        for (int lb = 0; lb < LANES; lb += CH) {
          double zv[CH], xv[CH];
#pragma unroll
          for (int l = 0; l < CH; l++)
            if (lb + l < nb) { zv[l] = __ldg(zq + (lb + l) * tk); xv[l] = rbuf[(lb + l) * 2 * (H + 1) + s]; }
#pragma unroll
          for (int l = 0; l < CH; l++)
            if (lb + l < nb) {
              xprev = (xv[l] - __ldg(a + k0 + lb + l) * xprev) * zv[l];
              rbuf[(lb + l) * 2 * (H + 1) + s] = xprev;
            }
        }
        carry[si] = xprev;
      }
    }
    __syncthreads();
    if (b < nbat - 1) {
      stage_out(k0, nb);
      __syncthreads();
    }
  }

  // ---------------- downward: back substitution (:1157-1163) + inverse transform ----------------
#pragma unroll 1
  for (int b = nbat - 1; b >= 0; b--) {
    const int k0 = b * LANES, nb = min(LANES, K - k0);
    const bool top = (b == nbat - 1);
    if (!top) {
      stage_in(k0, nb);
      __syncthreads();
    }
    const int l0 = top ? nb - 2 : nb - 1;   // the top level keeps x' (c(K) = 0); carry already holds it
#pragma unroll
    for (int si = 0; si < NSLOT; si++) {
      const int s = tid + si * NT;
      if (s < N) {
        const double *zq = ztp + ((s + 1) >> 1) + (long long)k0 * tk;
        double xnext = carry[si];
#pragma unroll 1
        for (int lb = LANES - CH; lb >= 0; lb -= CH) {
          double zv[CH], xv[CH];
#pragma unroll
          for (int l = 0; l < CH; l++)
            if (lb + l <= l0) { zv[l] = __ldg(zq + (lb + l) * tk); xv[l] = rbuf[(lb + l) * 2 * (H + 1) + s]; }
#pragma unroll
          for (int l = CH - 1; l >= 0; l--)
            if (lb + l <= l0) {
              xnext = xv[l] - (__ldg(c + k0 + lb + l) * zv[l]) * xnext;
              rbuf[(lb + l) * 2 * (H + 1) + s] = xnext;
            }
        }
        carry[si] = xnext;
      }
    }
    __syncthreads();
    xz_tile_inv<R1, R2, LANES>(buf, tw, fac, lane, j);
    __syncthreads();
    stage_out(k0, nb);
    __syncthreads();
  }
}

#undef XSA
#undef XRS

}  // namespace udg
