// poisson_fast.cuh — tuned Poisson kernels for power-of-two line lengths (64..1024):
//   * k_rfft_fast: batched real FFT with the reference's half-complex packing and 1/sqrt(n) scaling
//     (src/modpois.f90:478-490, :522-534 forward; :615-625, :669-679 inverse).  A length-n real line is
//     a length-h = n/2 complex FFT of z[m] = x[2m] + i x[2m+1] plus a split (forward) / merge
//     (inverse) pass.  h = R1*R2: every thread does one radix-R1 DFT entirely in registers, a twiddle,
//     ONE shared-memory exchange, then radix-R2 DFTs in registers (four-step FFT).  Lanes of a warp
//     run over the batch (32 neighbouring lines), so all index arithmetic is warp-uniform and every
//     shared-memory access is conflict-free; y lines (stride = row pitch) are read/written straight
//     from/to global memory, coalesced over the lanes; x lines (contiguous) are staged through a
//     padded shared-memory tile so global traffic stays coalesced along the line.
//   * k_zfactor / k_zsolve: the tridiagonal solve of solmpj (src/modpois.f90:1107-1166).  The Thomas
//     factors 1/(b_k + lambda - a_k d_{k-1}) depend only on the eigenvalue lambda = xrt(i) + yrt(j) and
//     the packed slots (2m, 2m+1) share eigenvalues (src/modpois.f90:100-107), so they are tabulated
//     once per distinct (lambda_x, lambda_y) pair (N/4 values).  The per-solve sweeps are then pure
//     FMA recurrences streamed with deep software prefetch, no divisions, no d scratch array.
#pragma once
#include "common.cuh"
#include "poisson_v1.cuh"

namespace udg {

// exp(-2 pi i k / 32) = (w32c(k), -w32s(k)); octant table, exactly symmetric
__host__ __device__ constexpr double w32q(int k) {
  return k == 0 ? 1.0 : k == 1 ? 0.9807852804032304 : k == 2 ? 0.9238795325112867 : k == 3 ? 0.8314696123025452
       : k == 4 ? 0.7071067811865476 : k == 5 ? 0.5555702330196022 : k == 6 ? 0.3826834323650898
       : k == 7 ? 0.1950903220161283 : 0.0;
}
__host__ __device__ constexpr double w32c(int k) {  // cos(2 pi k / 32), 0 <= k < 32
  return k <= 8 ? w32q(k) : k <= 16 ? -w32q(16 - k) : k <= 24 ? -w32q(k - 16) : w32q(32 - k);
}
__host__ __device__ constexpr double w32s(int k) {  // sin(2 pi k / 32)
  return k <= 8 ? w32q(8 - k) : k <= 16 ? w32q(k - 8) : k <= 24 ? -w32q(24 - k) : -w32q(k - 24);
}
template <int R>
__host__ __device__ constexpr int brev(int k) {
  int r = 0;
  for (int b = 1; b < R; b <<= 1) { r = (r << 1) | (k & 1); k >>= 1; }
  return r;
}

// in-register radix-2 DIF DFT of R points (R <= 32); result in bit-reversed order: X[k] = v[brev<R>(k)]
template <int R, bool INV>
__device__ __forceinline__ void dft_reg(double2 (&v)[R]) {
#pragma unroll
  for (int half = R / 2; half >= 1; half >>= 1) {
#pragma unroll
    for (int blk = 0; blk < R; blk += 2 * half) {
#pragma unroll
      for (int t = 0; t < half; t++) {
        const int i0 = blk + t, i1 = i0 + half;
        const double2 a = v[i0], b = v[i1];
        v[i0] = make_double2(a.x + b.x, a.y + b.y);
        const double dx = a.x - b.x, dy = a.y - b.y;
        const int e = t * (16 / half);  // twiddle exp(-/+ 2 pi i e / 32)
        if (e == 0) v[i1] = make_double2(dx, dy);
        else if (e == 8) v[i1] = INV ? make_double2(-dy, dx) : make_double2(dy, -dx);
        else {
          const double c = w32c(e), s = w32s(e);
          v[i1] = INV ? make_double2(dx * c - dy * s, dy * c + dx * s) : make_double2(dx * c + dy * s, dy * c - dx * s);
        }
      }
    }
  }
}

// Blocked ("wire format") addressing of one side of a transform for the multi-GPU transposes: the
// points of a line are split into nblk = n/blk consecutive runs, run d living in its own buffer
// base[d] (the block exchanged with rank d: local send/receive buffer or, with peer access, the
// receive buffer of GPU d itself).  Inside a block the LineDesc strides apply.  This is the pack /
// unpack (mem_split_* / mem_merge_*, 2decomp-fft/src/transpose_x_to_y.f90:275-327,385-437) fused into the
// FFT kernels' own loads and stores.
struct BlkDesc {
  double *base[8];
  int shift, mask;   // blk = 1 << shift points per block, mask = blk - 1
  // Output side only.  halo = 1: every block carries one extra column on each side (block pitch blk + 2 points): point q of
  // block d lands in column q + 1, and the first / last point of a block is ALSO stored as the last / first column of
  // the previous / next block (periodic over nblk blocks).  This is how the inverse x transform hands every slab the
  // two halo columns of p together with its own columns: bcp's exchange_halo_z (src/modboundary.f90:1344-1408) rides
  // on the transpose instead of being a separate exchange.
  int halo, nblk;
};

// Source of a FILL launch: the forward transform of the first pass evaluates the right-hand side of the Poisson equation
// itself while it loads its lines — fillps + bcpup (src/modpois.f90:911-973, src/modboundary.f90:1191-1255) fused into
// the transform: p = d/dx(up + um/c) + d/dy(vp + vm/c) + d/dz(wp + wm/c), pwp(kb) = pwp(ke+1) = 0, same expression and
// operation order as k_fillps, so the transformed data are the same bits as with the separate kernel.  The rhs array
// is never written nor read (16 B/cell less traffic), and in the slab solve the six input streams keep HBM busy while
// the transform's output is waiting for NVLink.
struct FillSrc {
  const double *up, *vp, *wp, *um, *vm, *wm;
  const double *dzfi;
  double rk3coefi, dxi, dyi;
  long long pi, pk;        // row and level pitch of the halo'd arrays (ih = jh = kh = 1)
  int imax, jmax, ktot;
  int xwrap;               // x unsplit: the +1 neighbour of the last column is the periodic image (else the halo column)
  int k0;                  // 0-based level of outer batch index 0 (k-chunks)
};
__device__ __forceinline__ double fill_rhs(const FillSrc &f, int i, int j, int k) {   // 1-based cell
  const int ip = (f.xwrap && i == f.imax) ? 1 : i + 1, jp = (j == f.jmax) ? 1 : j + 1;
  const long long t = (long long)i + f.pi * (j + 0ll) + f.pk * (k - 1), c = t + f.pk;     // offT / offF with halo 1
  const long long tx = (long long)ip + f.pi * (j + 0ll) + f.pk * (k - 1), ty = (long long)i + f.pi * (jp + 0ll) + f.pk * (k - 1);
  const double pu0 = f.up[t] + f.um[c] * f.rk3coefi;
  const double pu1 = f.up[tx] + f.um[tx + f.pk] * f.rk3coefi;
  const double pv0 = f.vp[t] + f.vm[c] * f.rk3coefi;
  const double pv1 = f.vp[ty] + f.vm[ty + f.pk] * f.rk3coefi;
  const double pw0 = (k == 1) ? 0.0 : f.wp[t] + f.wm[c] * f.rk3coefi;
  const double pw1 = (k == f.ktot) ? 0.0 : f.wp[t + f.pk] + f.wm[c + f.pk] * f.rk3coefi;
  return (pu1 - pu0) * f.dxi + (pv1 - pv0) * f.dyi + (pw1 - pw0) * f.dzfi[k];
}

template <int R1, int R2, int LANES, bool XDIR>
struct RfftCfg {
  static constexpr int H = R1 * R2, N = 2 * H, NT = LANES * R2;
  static constexpr int SMEM = XDIR ? LANES * (H + 1) * 16 : LANES * H * 16;
};

template <int R1, int R2, int LANES, bool XDIR, bool INV, bool IBLK = false, bool OBLK = false, bool FILL = false>
__global__ void __launch_bounds__(LANES *R2, (R1 <= 16 ? 2 : 1)) k_rfft_fast(const double2 *__restrict__ tw, const double *__restrict__ in, LineDesc di,
                                                         double *__restrict__ out, LineDesc dd, double fac,
                                                         BlkDesc ib = BlkDesc(), BlkDesc ob = BlkDesc(), FillSrc fs = FillSrc()) {
  static_assert(!FILL || (!INV && !IBLK), "FILL: forward transform reading the tendencies");
  constexpr int H = R1 * R2, N = 2 * H, NT = LANES * R2;
  constexpr int NK1 = R1 / R2;            // pass-2 DFTs per thread
  constexpr int NPAIR = (H / 2) / R2 + 1;  // split/merge pairs per thread (k = j, j+R2, ... <= H/2)
  static_assert(R1 % R2 == 0, "R1 must be a multiple of R2");
  extern __shared__ double2 buf[];
  double *rbuf = reinterpret_cast<double *>(buf);
  const int lane = threadIdx.x, j = threadIdx.y, tid = j * LANES + lane;
  const int b0 = blockIdx.x * LANES;
  const int nb = min(LANES, di.nb1 - b0);
  const bool act = lane < nb;
  const int kb = di.rev ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const long long ibase = (long long)kb * di.s2 + (long long)b0 * di.s1;
  const long long obase = (long long)kb * dd.s2 + (long long)b0 * dd.s1;
  // element addresses: (line offset within the CTA's batch, point index)
  auto IN = [&](long long loff, int pt) -> const double * {
    return IBLK ? ib.base[pt >> ib.shift] + ibase + loff + (long long)(pt & ib.mask) * di.sp : in + ibase + loff + (long long)pt * di.sp;
  };
  auto OUT = [&](long long loff, int pt) -> double * {
    return OBLK ? ob.base[pt >> ob.shift] + obase + loff + (long long)((pt & ob.mask) + ob.halo) * dd.sp : out + obase + loff + (long long)pt * dd.sp;
  };
  // halo-carrying blocks: duplicate an edge point into the neighbouring block's halo column
  auto OUT_EDGE = [&](long long loff, int pt, double val) {
    if (!OBLK || !ob.halo) return;
    const int q = pt & ob.mask, d = pt >> ob.shift;
    if (q == 0) { const int dl = d == 0 ? ob.nblk - 1 : d - 1; ob.base[dl][obase + loff + (long long)(ob.mask + 2) * dd.sp] = val; }
    if (q == ob.mask) { const int dr = d == ob.nblk - 1 ? 0 : d + 1; ob.base[dr][obase + loff] = val; }
  };
  // FILL: value of point pt of line b of this CTA's batch = rhs(i, j, k) evaluated on the fly
  auto RHS = [&](int b, int pt) -> double {
    const int k = fs.k0 + kb + 1;
    return XDIR ? fill_rhs(fs, pt + 1, b0 + b + 1, k) : fill_rhs(fs, b0 + b + 1, pt + 1, k);
  };
  // x lines: global -> padded shared tile.  All of a thread's loads are issued before the first use (R1
  // independent 16-byte requests in flight per thread); 128-bit accesses when the lines are 16-byte aligned.
  constexpr int NIT = (LANES * H) / NT;   // = R1
  const bool al_in = !IBLK ? ((((size_t)(in + ibase)) & 15) == 0 && ((di.s1 | di.s2) & 1) == 0)
                           : ((((size_t)ib.base[0]) & 15) == 0 && ((di.s1 | di.s2 | ibase) & 1) == 0);
  const bool al_out = !OBLK ? ((((size_t)(out + obase)) & 15) == 0 && ((dd.s1 | dd.s2) & 1) == 0)
                            : ((((size_t)ob.base[0]) & 15) == 0 && ((dd.s1 | dd.s2 | obase) & 1) == 0 && !ob.halo);
  constexpr int SCH = NIT > 8 ? 8 : NIT;   // requests in flight per thread and staging round
  auto stage_in = [&]() {
#pragma unroll
    for (int i0 = 0; i0 < NIT; i0 += SCH) {
      double2 st[SCH];
#pragma unroll
      for (int it = 0; it < SCH; it++) {
        const int idx = tid + (i0 + it) * NT;
        if (idx < nb * H) {
          const int b = idx / H, m = idx - b * H;
          if (FILL) st[it] = make_double2(RHS(b, 2 * m), RHS(b, 2 * m + 1));
          else {
            const double *q = IN((long long)b * di.s1, 2 * m);
            st[it] = al_in ? *reinterpret_cast<const double2 *>(q) : make_double2(q[0], q[1]);
          }
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
  auto stage_out = [&]() {
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
          double *q = OUT((long long)b * dd.s1, 2 * m);
          if (al_out) *reinterpret_cast<double2 *>(q) = st[it];
          else { q[0] = st[it].x; q[1] = st[it].y; }
          OUT_EDGE((long long)b * dd.s1, 2 * m, st[it].x);
          OUT_EDGE((long long)b * dd.s1, 2 * m + 1, st[it].y);
        }
      }
    }
  };
#define SA(p) (XDIR ? (lane * (H + 1) + (p)) : ((p)*LANES + lane))
#define RS(r) (lane * 2 * (H + 1) + (r))  /* real slot r of this lane's line in the x staging tile */

  double2 v[R1];
  if (!INV) {
    if (XDIR) {
      stage_in();
      __syncthreads();
#pragma unroll
      for (int q = 0; q < R1; q++) v[q] = buf[SA(j + R2 * q)];
      __syncthreads();
    } else if (act) {
      const long long lo = (long long)lane * di.s1;
#pragma unroll
      for (int q = 0; q < R1; q++) {
        const int m = j + R2 * q;
        v[q] = FILL ? make_double2(RHS(lane, 2 * m), RHS(lane, 2 * m + 1)) : make_double2(*IN(lo, 2 * m), *IN(lo, 2 * m + 1));
      }
    }
  } else {
    // merge: Z[k] = A + T, Z[h-k] = conj(A - T), A = Xk + conj(Xhk), T = i conj(w^k) (Xk - conj(Xhk))
    if (XDIR) {
      stage_in();
      __syncthreads();
    }
    double2 zk[NPAIR], zh[NPAIR];
    // y lines: issue every global load of the thread first (4 NPAIR independent requests in flight), then merge
    double gx0[NPAIR], gx1[NPAIR], gy0[NPAIR], gy1[NPAIR];
    if (!XDIR) {
      const long long lo = (long long)lane * di.s1;
#pragma unroll
      for (int t = 0; t < NPAIR; t++) {
        const int k = j + R2 * t;
        gx0[t] = gx1[t] = gy0[t] = gy1[t] = 0.;
        if (k <= H / 2 && act) {
          if (k == 0) { gx0[t] = *IN(lo, 0); gy0[t] = *IN(lo, N - 1); }
          else {
            gx0[t] = *IN(lo, 2 * k - 1); gx1[t] = *IN(lo, 2 * k);
            gy0[t] = *IN(lo, 2 * (H - k) - 1); gy1[t] = *IN(lo, 2 * (H - k));
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < NPAIR; t++) {
      const int k = j + R2 * t;
      zk[t] = zh[t] = make_double2(0., 0.);
      if (k <= H / 2 && act) {
        double x0, x1, y0, y1;
        if (k == 0) {
          if (XDIR) { x0 = rbuf[RS(0)]; y0 = rbuf[RS(N - 1)]; }
          else { x0 = gx0[t]; y0 = gy0[t]; }
          zk[t] = make_double2(x0 + y0, x0 - y0);
        } else {
          if (XDIR) { x0 = rbuf[RS(2 * k - 1)]; x1 = rbuf[RS(2 * k)]; y0 = rbuf[RS(2 * (H - k) - 1)]; y1 = rbuf[RS(2 * (H - k))]; }
          else { x0 = gx0[t]; x1 = gx1[t]; y0 = gy0[t]; y1 = gy1[t]; }
          const double2 A = make_double2(x0 + y0, x1 - y1), Bv = make_double2(x0 - y0, x1 + y1);
          const double2 w = tw[k];
          const double2 T = cmul(make_double2(w.y, w.x), Bv);
          zk[t] = cadd(A, T);
          zh[t] = cconj(csub(A, T));
        }
      }
    }
    if (XDIR) __syncthreads();
#pragma unroll
    for (int t = 0; t < NPAIR; t++) {
      const int k = j + R2 * t;
      if (k <= H / 2) {
        if (k != 0 && k != H - k) buf[SA(H - k)] = zh[t];
        buf[SA(k)] = zk[t];
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < R1; q++) v[q] = buf[SA(j + R2 * q)];
    __syncthreads();
  }

  // ---- pass 1: radix-R1 over q, twiddle W_h^{j k1}, exchange ----
  dft_reg<R1, INV>(v);
#pragma unroll
  for (int k1 = 0; k1 < R1; k1++) {
    double2 y = v[brev<R1>(k1)];
    if (k1 != 0) {
      double2 w = tw[2 * j * k1];  // W_h^{j k1} = exp(-2 pi i 2 j k1 / n);  j*k1 < h so 2 j k1 < n
      if (INV) w.y = -w.y;
      y = cmul(y, w);
    }
    buf[SA(k1 * R2 + j)] = y;
  }
  __syncthreads();
  // ---- pass 2: radix-R2 over j for k1 = j + R2*t ----
  double2 u[NK1][R2];
#pragma unroll
  for (int t = 0; t < NK1; t++) {
    const int k1 = j + R2 * t;
#pragma unroll
    for (int jj = 0; jj < R2; jj++) u[t][jj] = buf[SA(k1 * R2 + jj)];
    dft_reg<R2, INV>(u[t]);
  }

  if (!INV) {
    __syncthreads();
    // natural order: Z[k1 + R1 k2]
#pragma unroll
    for (int t = 0; t < NK1; t++)
#pragma unroll
      for (int k2 = 0; k2 < R2; k2++) buf[SA(j + R2 * t + R1 * k2)] = u[t][brev<R2>(k2)];
    __syncthreads();
    // split: X[k] = E + T, X[h-k] = conj(E - T), E = (Zk + conj Zhk)/2, T = -i/2 w^k (Zk - conj Zhk)
    double2 xk[NPAIR], xh[NPAIR];
#pragma unroll
    for (int t = 0; t < NPAIR; t++) {
      const int k = j + R2 * t;
      xk[t] = xh[t] = make_double2(0., 0.);
      if (k <= H / 2) {
        const double2 Zk = buf[SA(k)];
        if (k == 0) {
          xk[t] = make_double2((Zk.x + Zk.y) * fac, (Zk.x - Zk.y) * fac);  // (X0, Xh) both real
        } else {
          const double2 Zc = cconj(buf[SA(H - k)]);
          const double2 E = make_double2(0.5 * (Zk.x + Zc.x), 0.5 * (Zk.y + Zc.y));
          const double2 D = csub(Zk, Zc);
          const double2 w = tw[k];
          const double2 T = cmul(make_double2(0.5 * w.y, -0.5 * w.x), D);
          const double2 a = cadd(E, T), b = cconj(csub(E, T));
          xk[t] = make_double2(a.x * fac, a.y * fac);
          xh[t] = make_double2(b.x * fac, b.y * fac);
        }
      }
    }
    if (XDIR) {
      __syncthreads();
#pragma unroll
      for (int t = 0; t < NPAIR; t++) {
        const int k = j + R2 * t;
        if (k <= H / 2) {
          if (k == 0) { rbuf[RS(0)] = xk[t].x; rbuf[RS(N - 1)] = xk[t].y; }
          else {
            rbuf[RS(2 * k - 1)] = xk[t].x; rbuf[RS(2 * k)] = xk[t].y;
            rbuf[RS(2 * (H - k) - 1)] = xh[t].x; rbuf[RS(2 * (H - k))] = xh[t].y;
          }
        }
      }
      __syncthreads();
      stage_out();
    } else if (act) {
      const long long lo = (long long)lane * dd.s1;
#pragma unroll
      for (int t = 0; t < NPAIR; t++) {
        const int k = j + R2 * t;
        if (k <= H / 2) {
          if (k == 0) { *OUT(lo, 0) = xk[t].x; *OUT(lo, N - 1) = xk[t].y; }
          else {
            *OUT(lo, 2 * k - 1) = xk[t].x; *OUT(lo, 2 * k) = xk[t].y;
            *OUT(lo, 2 * (H - k) - 1) = xh[t].x; *OUT(lo, 2 * (H - k)) = xh[t].y;
          }
        }
      }
    }
  } else {
    // inverse: z[m], m = k1 + R1 k2 -> x[2m] = Re, x[2m+1] = Im, times fac
    if (XDIR) {
      __syncthreads();
#pragma unroll
      for (int t = 0; t < NK1; t++)
#pragma unroll
        for (int k2 = 0; k2 < R2; k2++) {
          const double2 z = u[t][brev<R2>(k2)];
          buf[SA(j + R2 * t + R1 * k2)] = make_double2(z.x * fac, z.y * fac);
        }
      __syncthreads();
      stage_out();
    } else if (act) {
      const long long lo = (long long)lane * dd.s1;
#pragma unroll
      for (int t = 0; t < NK1; t++)
#pragma unroll
        for (int k2 = 0; k2 < R2; k2++) {
          const int m = j + R2 * t + R1 * k2;
          const double2 z = u[t][brev<R2>(k2)];
          *OUT(lo, 2 * m) = z.x * fac;
          *OUT(lo, 2 * m + 1) = z.y * fac;
        }
    }
  }
#undef SA
#undef RS
}

// ---------------------------------------------------------------------------------------------
// Thomas factor table.  zt[(k*nyh + jy)*nxh + ix] = 1 / (bb_k - a_k d_{k-1}), d_k = c_k z_k, for the
// eigenvalue pair lambda = xd[ix] + yd[jy] (distinct eigenvalues: ix = 0..itot/2, jy = 0..jtot/2).
// Same recurrence and special cases as src/modpois.f90:1120-1155 with bxyzrt of :196-220.
__global__ void k_zfactor(int nxh, int nyh, int K, const double *__restrict__ xd, const double *__restrict__ yd,
                          const double *__restrict__ a, const double *__restrict__ b, const double *__restrict__ c,
                          double b_top_D, double *__restrict__ zt) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x, jy = blockIdx.y;
  if (ix >= nxh) return;
  const double lam = xd[ix] + yd[jy];
  const long long sk = (long long)nxh * nyh;
  long long q = (long long)jy * nxh + ix;
  double bb = (lam == 0. && K == 1) ? b_top_D : b[0] + lam;
  double z = 1. / bb, d = c[0] * z;
  zt[q] = z;
  for (int k = 1; k < K; k++) {
    q += sk;
    bb = (k == K - 1 && lam == 0.) ? b_top_D : b[k] + lam;
    z = 1. / (bb - a[k] * d);
    d = c[k] * z;
    zt[q] = z;
  }
}

// One thread per (i,j) column of the halo-free spectral array x(imax,jmax,K); forward then backward
// sweep in place.  The recurrence is a dependent FMA chain, so memory latency is hidden by explicit
// double buffering: while batch A (ZU levels) runs through the chain, the loads of batch B are in flight.
template <int ZU>
__global__ void __launch_bounds__(128) k_zsolve(Geo g, int nxh, int nyh, double *__restrict__ x, const double *__restrict__ zt,
                                                const double *__restrict__ a, const double *__restrict__ c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= g.imax) return;
  const int K = g.ktot;
  const int ig = g.i0g + i, jg = g.j0g + j;       // global slot (0-based) -> distinct-eigenvalue index
  const int ix = g.xalt ? (ig == 0 ? 0 : ig == 1 ? nxh - 1 : ig >> 1) : (ig + 1) >> 1, jy = (jg + 1) >> 1;
  const long long sk = (long long)g.imax * g.jmax, tk = (long long)nxh * nyh;
  double *xp = x + (long long)i + (long long)g.imax * j;
  const double *zp = zt + (long long)jy * nxh + ix;
  double xa[ZU], za[ZU], xb[ZU], zb[ZU];
  const int nb = K / ZU;
#define ZLOADF(X, Z, k0) _Pragma("unroll") for (int u = 0; u < ZU; u++) { X[u] = xp[((k0) + u) * sk]; Z[u] = __ldg(zp + ((k0) + u) * tk); }
#define ZRUNF(X, Z, k0) _Pragma("unroll") for (int u = 0; u < ZU; u++) { xprev = (X[u] - __ldg(a + (k0) + u) * xprev) * Z[u]; xp[((k0) + u) * sk] = xprev; }
  double xprev = 0.;
  if (nb > 0) ZLOADF(xa, za, 0)
  for (int b = 0; b < nb; b += 2) {
    if (b + 1 < nb) ZLOADF(xb, zb, (b + 1) * ZU)
    ZRUNF(xa, za, b * ZU)
    if (b + 1 < nb) {
      if (b + 2 < nb) ZLOADF(xa, za, (b + 2) * ZU)
      ZRUNF(xb, zb, (b + 1) * ZU)
    }
  }
  for (int k = nb * ZU; k < K; k++) {
    xprev = (xp[k * sk] - __ldg(a + k) * xprev) * __ldg(zp + k * tk);
    xp[k * sk] = xprev;
  }
#undef ZLOADF
#undef ZRUNF
  // backward: x_k = x'_k - d_k x_{k+1}, d_k = c_k z_k  (c_{K-1} = 0), levels K-2 .. 0
  double xnext = xprev;
  const int nbb = (K - 1) / ZU;
#define ZLOADB(X, Z, k0) _Pragma("unroll") for (int u = 0; u < ZU; u++) { X[u] = xp[((k0) - u) * sk]; Z[u] = __ldg(zp + ((k0) - u) * tk); }
#define ZRUNB(X, Z, k0) _Pragma("unroll") for (int u = 0; u < ZU; u++) { xnext = X[u] - (__ldg(c + (k0) - u) * Z[u]) * xnext; xp[((k0) - u) * sk] = xnext; }
  if (nbb > 0) ZLOADB(xa, za, K - 2)
  for (int b = 0; b < nbb; b += 2) {
    if (b + 1 < nbb) ZLOADB(xb, zb, K - 2 - (b + 1) * ZU)
    ZRUNB(xa, za, K - 2 - b * ZU)
    if (b + 1 < nbb) {
      if (b + 2 < nbb) ZLOADB(xa, za, K - 2 - (b + 2) * ZU)
      ZRUNB(xb, zb, K - 2 - (b + 1) * ZU)
    }
  }
  for (int k = K - 2 - nbb * ZU; k >= 0; k--) {
    xnext = xp[k * sk] - (__ldg(c + k) * __ldg(zp + k * tk)) * xnext;
    xp[k * sk] = xnext;
  }
#undef ZLOADB
#undef ZRUNB
}

}  // namespace udg
