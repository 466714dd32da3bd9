// poisson_xline.cuh — x-direction real FFT of the Poisson solve with the threads of a line sitting NEXT TO each other.
//
// Same transform as k_rfft_fast<.., XDIR = true, ..> (poisson_fast.cuh: four-step FFT h = R1 * R2 of z[m] = x[2m] + i x[2m+1],
// split / merge pass, the reference's half-complex packing [Re0, Re1, Im1, .., Re(n/2)] and 1/sqrt(n) scaling,
// src/modpois.f90:478-490, 669-679), same operations in the same order on every value — the results are the same bits —
// but another thread layout.  k_rfft_fast maps the 32 lanes of a warp to 32 different lines, which suits the y direction
// (a warp then reads 32 neighbouring i of one j: coalesced) and forces the x direction, whose lines are contiguous in
// memory, through a shared-memory staging tile on the way in and on the way out: three shared-memory round trips and
// seven block-wide barriers more than the y pass, 78 / 89 us against 55 / 59 us at 256^3 (profiles/r2_launches_n1.csv;
// ncu: MIO / short-scoreboard stalls).  Here the R2 threads of a line are consecutive lanes:
//   * pass-1 input z[j + R2 q] (thread j, q = 0 .. R1-1) is R2 consecutive 16-byte words per q: the global loads are
//     coalesced as they stand, nothing is staged;
//   * both exchanges of the four-step scheme stay inside one warp (R2 <= 32 lanes): __syncwarp instead of __syncthreads,
//     no block-wide barrier at all;
//   * the results leave in 8 / 16-byte stores straight from registers: consecutive lanes own consecutive output words.
// The tile of a line in shared memory is padded by one word per R2 (pad(p) = p + p / R2) so that the pass-2 reads
// (thread stride R2 words) fall into different banks.
//
// Spectral layout BETWEEN the two x passes.  The reference's packing [Re0, Re1, Im1, .., Re(n/2)] puts every complex
// coefficient across a 16-byte boundary: 36 eight-byte stores / loads per thread, half a sector each (ncu: LSU data pipe
// 69 % of its peak, the limiter of the first version of this kernel).  The work array is ours between the forward and
// the inverse x pass, so the line is kept as aligned pairs instead: slots (0, 1) = (X0, X(n/2)) — both real — and slots
// (2k, 2k+1) = (Re Xk, Im Xk), k = 1 .. n/2-1.  The y transforms treat every x slot as an independent column, and the z
// solve only needs the slot -> eigenvalue map (Geo::xalt: slot 0 -> 0, slot 1 -> n/2, slots 2k, 2k+1 -> k; packed:
// slot s -> (s+1)/2), so every number goes through the same arithmetic as before and the pressure is the same bits.
#pragma once
#include "common.cuh"
#include "poisson_fast.cuh"

namespace udg {

template <int R1, int R2>
struct XlineCfg {
  static constexpr int H = R1 * R2, N = 2 * H, NT = 256, LPB = NT / R2, PITCH = H + H / R2 + 1;
  static constexpr int SMEM = LPB * PITCH * 16;
};

template <int R1, int R2, bool INV>
__global__ void __launch_bounds__(256, (R1 >= 32 ? 1 : 2)) k_rfft_xline(const double2 *__restrict__ tw, const double *__restrict__ in, LineDesc di,
                                                       double *__restrict__ out, LineDesc dd, double fac) {
  using C = XlineCfg<R1, R2>;
  constexpr int H = C::H, N = C::N, LPB = C::LPB, PITCH = C::PITCH;
  constexpr int NK1 = R1 / R2;             // pass-2 DFTs per thread
  constexpr int NPAIR = (H / 2) / R2 + 1;  // split / merge pairs per thread (k = j, j + R2, .. <= H/2)
  static_assert(R1 % R2 == 0 && 32 % R2 == 0, "a line's R2 threads must be lanes of one warp");
  extern __shared__ double2 xl_buf[];
  const int tid = threadIdx.x, j = tid % R2, l = tid / R2;
  const int line = blockIdx.x * LPB + l;
  const bool act = line < di.nb1;
  const int kb = di.rev ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const double *pin = in + (long long)kb * di.s2 + (long long)(act ? line : 0) * di.s1;    // di.sp == dd.sp == 1: x lines
  double *pout = out + (long long)kb * dd.s2 + (long long)(act ? line : 0) * dd.s1;
  const bool al_in = (((size_t)pin) & 15) == 0, al_out = (((size_t)pout) & 15) == 0;
  double2 *buf = xl_buf + l * PITCH;
#define SX(p) ((p) + (p) / R2)

  double2 v[R1];
  if (!INV) {
#pragma unroll
    for (int q = 0; q < R1; q++) {
      const double *s = pin + 2 * (j + R2 * q);
      v[q] = !act ? make_double2(0., 0.) : al_in ? *reinterpret_cast<const double2 *>(s) : make_double2(s[0], s[1]);
    }
  } else {
    // merge: Z[k] = A + T, Z[h-k] = conj(A - T), A = Xk + conj(Xhk), T = i conj(w^k) (Xk - conj(Xhk)); all loads first
    double gx0[NPAIR], gx1[NPAIR], gy0[NPAIR], gy1[NPAIR];
#pragma unroll
    for (int t = 0; t < NPAIR; t++) {
      const int k = j + R2 * t;
      gx0[t] = gx1[t] = gy0[t] = gy1[t] = 0.;
      if (k <= H / 2 && act) {
        const double *sa = pin + 2 * k, *sb = pin + 2 * (H - k);     // aligned pairs (see the header): (X0, Xh) / (Re Xk, Im Xk)
        if (al_in) {
          const double2 a = *reinterpret_cast<const double2 *>(sa);
          gx0[t] = a.x; gx1[t] = a.y;
          if (k == 0) gy0[t] = a.y;
          else { const double2 b = *reinterpret_cast<const double2 *>(sb); gy0[t] = b.x; gy1[t] = b.y; }
        } else {
          gx0[t] = sa[0]; gx1[t] = sa[1];
          if (k == 0) gy0[t] = sa[1];
          else { gy0[t] = sb[0]; gy1[t] = sb[1]; }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < NPAIR; t++) {
      const int k = j + R2 * t;
      if (k <= H / 2) {
        double2 zk, zh = make_double2(0., 0.);
        if (k == 0) zk = make_double2(gx0[t] + gy0[t], gx0[t] - gy0[t]);
        else {
          const double2 A = make_double2(gx0[t] + gy0[t], gx1[t] - gy1[t]), Bv = make_double2(gx0[t] - gy0[t], gx1[t] + gy1[t]);
          const double2 w = tw[k];
          const double2 T = cmul(make_double2(w.y, w.x), Bv);
          zk = cadd(A, T);
          zh = cconj(csub(A, T));
        }
        if (k != 0 && k != H - k) buf[SX(H - k)] = zh;
        buf[SX(k)] = zk;
      }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < R1; q++) v[q] = buf[SX(j + R2 * q)];
    __syncwarp();
  }

  // ---- pass 1: radix-R1 over q, twiddle W_h^{j k1}, exchange inside the warp ----
  dft_reg<R1, INV>(v);
#pragma unroll
  for (int k1 = 0; k1 < R1; k1++) {
    double2 y = v[brev<R1>(k1)];
    if (k1 != 0) {
      double2 w = tw[2 * j * k1];
      if (INV) w.y = -w.y;
      y = cmul(y, w);
    }
    buf[SX(k1 * R2 + j)] = y;
  }
  __syncwarp();
  // ---- pass 2: radix-R2 over j for k1 = j + R2 t ----
  double2 u[NK1][R2];
#pragma unroll
  for (int t = 0; t < NK1; t++) {
    const int k1 = j + R2 * t;
#pragma unroll
    for (int jj = 0; jj < R2; jj++) u[t][jj] = buf[SX(k1 * R2 + jj)];
    dft_reg<R2, INV>(u[t]);
  }

  if (!INV) {
    __syncwarp();
    // natural order: Z[k1 + R1 k2]
#pragma unroll
    for (int t = 0; t < NK1; t++)
#pragma unroll
      for (int k2 = 0; k2 < R2; k2++) buf[SX(j + R2 * t + R1 * k2)] = u[t][brev<R2>(k2)];
    __syncwarp();
    // split: X[k] = E + T, X[h-k] = conj(E - T), E = (Zk + conj Zhk)/2, T = -i/2 w^k (Zk - conj Zhk)
#pragma unroll
    for (int t = 0; t < NPAIR; t++) {
      const int k = j + R2 * t;
      if (k <= H / 2) {
        const double2 Zk = buf[SX(k)];
        if (k == 0) {
          if (act) {   // (X0, Xh) both real: slots 0, 1
            const double2 o = make_double2((Zk.x + Zk.y) * fac, (Zk.x - Zk.y) * fac);
            if (al_out) *reinterpret_cast<double2 *>(pout) = o; else { pout[0] = o.x; pout[1] = o.y; }
          }
        } else {
          const double2 Zc = cconj(buf[SX(H - k)]);
          const double2 E = make_double2(0.5 * (Zk.x + Zc.x), 0.5 * (Zk.y + Zc.y));
          const double2 D = csub(Zk, Zc);
          const double2 w = tw[k];
          const double2 T = cmul(make_double2(0.5 * w.y, -0.5 * w.x), D);
          const double2 a = cadd(E, T), b = cconj(csub(E, T));
          if (act) {
            const double2 oa = make_double2(a.x * fac, a.y * fac), ob = make_double2(b.x * fac, b.y * fac);
            double *da = pout + 2 * k, *db = pout + 2 * (H - k);
            if (al_out) { *reinterpret_cast<double2 *>(da) = oa; *reinterpret_cast<double2 *>(db) = ob; }
            else { da[0] = oa.x; da[1] = oa.y; db[0] = ob.x; db[1] = ob.y; }
          }
        }
      }
    }
  } else if (act) {
    // inverse: z[m], m = k1 + R1 k2 -> x[2m] = Re, x[2m+1] = Im, times fac
#pragma unroll
    for (int t = 0; t < NK1; t++)
#pragma unroll
      for (int k2 = 0; k2 < R2; k2++) {
        const int m = j + R2 * t + R1 * k2;
        const double2 z = u[t][brev<R2>(k2)];
        double *d = pout + 2 * m;
        if (al_out) *reinterpret_cast<double2 *>(d) = make_double2(z.x * fac, z.y * fac);
        else { d[0] = z.x * fac; d[1] = z.y * fac; }
      }
  }
#undef SX
}

}  // namespace udg
