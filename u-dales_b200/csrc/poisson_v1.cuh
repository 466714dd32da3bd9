// poisson_v1.cuh — general-size batched real FFT (any n = 2*h with h = product of primes <= 13)
// and the streaming Thomas solve.  This is the always-correct path; the tuned power-of-two
// register FFTs and the fused z-solve live in poisson_fast.cuh.
//
// Transform conventions = what the reference gets from FFTW r2c/c2r plus its own packing and
// 1/sqrt(n) scaling (src/modpois.f90:478-490 forward, :669-679 inverse):
//   forward: y = [Re X0, Re X1, Im X1, ..., Re X(h-1), Im X(h-1), Re Xh] / sqrt(n),  X = sum x e^{-2 pi i jk/n}
//   inverse: x = c2r_unnormalised(X) / sqrt(n)
// A length-n real transform is done as a length-h complex Stockham FFT of z[m] = x[2m] + i x[2m+1]
// with a split/merge pass; the complex FFT is an in-place mixed-radix decimation-in-frequency
// transform whose digit-reversed output order is undone by index (fft_perm) when it is consumed.
#pragma once
#include "common.cuh"

namespace udg {

struct LineDesc {
  long long sp;   // element stride between consecutive points of one line
  long long s1;   // element stride of the inner batch index (mapped to lanes)
  long long s2;   // element stride of the outer batch index (mapped to blockIdx.y)
  int nb1, nb2;   // batch extents
  int rev;        // 1: walk the outer batch index downwards (blockIdx.y = 0 handles nb2-1).  Consecutive passes over the
                  // same array alternate direction so that a pass starts on the part its predecessor wrote last and
                  // that is still in L2 (the array is larger than L2: same-direction streaming would miss everywhere)
};

struct FftPlan {
  int n, h;         // real length, complex length n/2
  int nst;          // number of Stockham stages
  int radix[24];
  const double2 *tw;  // tw[q] = exp(-2 pi i q / n), q < n   (host-computed in fp64)
  double fac;       // 1/sqrt(n)
};

constexpr int FFT_B = 16;        // lines per CTA (lanes of the batch dimension)
constexpr int FFT_BP = FFT_B + 1;  // padded pitch of the smem tile (complex elements)
constexpr int FFT_TY = 16;

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// position of DFT bin k after the in-place DIF stages (digit reversal in the mixed radix system)
__device__ __forceinline__ int fft_perm(const FftPlan &pl, int k) {
  int L = pl.h, pos = 0;
  for (int s = 0; s < pl.nst; s++) {
    const int r = pl.radix[s], M = L / r;
    pos += ((k / (pl.h / L)) % r) * M;
    L = M;
  }
  return pos;
}

// One CTA: FFT_B lines (inner batch index b1 = blockIdx.x*FFT_B + lane) of outer index blockIdx.y.
// CONTIG: the points of a line are contiguous in memory (x lines) -> global accesses run along the
// points; otherwise (y lines, sp = row pitch, s1 = 1) they run along the lanes.
template <bool CONTIG>
__global__ void __launch_bounds__(FFT_B *FFT_TY) k_rfft(FftPlan pl, int inverse, const double *__restrict__ in, LineDesc di,
                                                        double *__restrict__ out, LineDesc dd) {
  extern __shared__ double2 buf[];  // [h][FFT_BP]
  const int h = pl.h, n = pl.n;
  const int lane = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * FFT_B + lane, nthr = FFT_B * FFT_TY;
  const int b1_0 = blockIdx.x * FFT_B;
  const int nb = min(FFT_B, di.nb1 - b1_0);
  const long long ibase = (long long)blockIdx.y * di.s2 + (long long)b1_0 * di.s1;
  const long long obase = (long long)blockIdx.y * dd.s2 + (long long)b1_0 * dd.s1;
  const double2 *__restrict__ tw = pl.tw;

  // ---- phase 1: load (+ merge for the inverse) ----
  if (!inverse) {
    if (CONTIG) {
      for (int idx = tid; idx < nb * h; idx += nthr) {
        const int b = idx / h, m = idx - b * h;
        const double *q = in + ibase + b * di.s1 + 2 * m * di.sp;
        buf[m * FFT_BP + b] = make_double2(q[0], q[di.sp]);
      }
    } else if (lane < nb) {
      for (int m = ty; m < h; m += FFT_TY) {
        const double *q = in + ibase + lane * di.s1 + 2 * m * di.sp;
        buf[m * FFT_BP + lane] = make_double2(q[0], q[di.sp]);
      }
    }
  } else {
    // Z[k] = A + T, Z[h-k] = conj(A - T), A = Xk + conj(Xhk), T = i conj(w^k) (Xk - conj(Xhk))
    const int npair = h / 2 + 1;  // k = 0 .. h/2
    if (CONTIG) {
      for (int idx = tid; idx < nb * npair; idx += nthr) {
        const int b = idx / npair, k = idx - b * npair;
        const double *q = in + ibase + b * di.s1;
        if (k == 0) {
          const double x0 = q[0], xh = q[(long long)(n - 1) * di.sp];
          buf[b] = make_double2(x0 + xh, x0 - xh);
        } else {
          const double2 Xk = make_double2(q[(2 * k - 1) * di.sp], q[(2 * k) * di.sp]);
          const double2 Xh = make_double2(q[(2 * (h - k) - 1) * di.sp], q[(2 * (h - k)) * di.sp]);
          const double2 A = cadd(Xk, cconj(Xh)), Bv = csub(Xk, cconj(Xh));
          const double2 w = tw[k];                       // exp(-2 pi i k/n)
          const double2 T = cmul(make_double2(w.y, w.x), Bv);  // i*conj(w) = (w.y, w.x)
          buf[k * FFT_BP + b] = cadd(A, T);
          buf[(h - k) * FFT_BP + b] = cconj(csub(A, T));
        }
      }
    } else if (lane < nb) {
      const double *q = in + ibase + lane * di.s1;
      for (int k = ty; k < npair; k += FFT_TY) {
        if (k == 0) {
          const double x0 = q[0], xh = q[(long long)(n - 1) * di.sp];
          buf[lane] = make_double2(x0 + xh, x0 - xh);
        } else {
          const double2 Xk = make_double2(q[(2 * k - 1) * di.sp], q[(2 * k) * di.sp]);
          const double2 Xh = make_double2(q[(2 * (h - k) - 1) * di.sp], q[(2 * (h - k)) * di.sp]);
          const double2 A = cadd(Xk, cconj(Xh)), Bv = csub(Xk, cconj(Xh));
          const double2 w = tw[k];
          const double2 T = cmul(make_double2(w.y, w.x), Bv);
          buf[k * FFT_BP + lane] = cadd(A, T);
          buf[(h - k) * FFT_BP + lane] = cconj(csub(A, T));
        }
      }
    }
  }
  __syncthreads();

  // ---- phase 2: in-place decimation-in-frequency stages, lanes = batch, generic radix r ----
  // stage: block length L, M = L/r; butterfly (B0, j): inputs B0 + j + q*M, outputs to the same
  // slots: y_p = W_L^{p j} * sum_q x_q W_r^{p q}.  Result ends up digit-reversed (fft_perm).
  {
    int L = h;
    for (int s = 0; s < pl.nst; s++) {
      const int r = pl.radix[s];
      const int M = L / r;
      const int nbf = h / r;
      const int tL = n / L;    // W_L^{e} = tw[(e * tL) mod n]
      const int tr = n / r;    // W_r^{e} = tw[(e mod r) * tr]
      if (lane < nb) {
        for (int t = ty; t < nbf; t += FFT_TY) {
          const int blk = t / M, j = t - blk * M;
          double2 *x = buf + (long long)(blk * L + j) * FFT_BP + lane;
          const long long qs = (long long)M * FFT_BP;
          if (r == 2) {
            const double2 a = x[0], b = x[qs];
            x[0] = cadd(a, b);
            double2 d = csub(a, b);
            if (j) { double2 w = tw[j * tL]; if (inverse) w.y = -w.y; d = cmul(d, w); }
            x[qs] = d;
          } else if (r == 4) {
            const double2 a = x[0], b = x[qs], c = x[2 * qs], d = x[3 * qs];
            const double2 t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = csub(b, d);
            // forward: -i*t3 = (t3.y, -t3.x); inverse: +i*t3 = (-t3.y, t3.x)
            const double2 it3 = inverse ? make_double2(-t3.y, t3.x) : make_double2(t3.y, -t3.x);
            double2 y0 = cadd(t0, t2), y1 = cadd(t1, it3), y2 = csub(t0, t2), y3 = csub(t1, it3);
            if (j) {
              double2 w1 = tw[j * tL], w2 = tw[(2 * j * tL) % n], w3 = tw[(int)(((long long)3 * j * tL) % n)];
              if (inverse) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
              y1 = cmul(y1, w1); y2 = cmul(y2, w2); y3 = cmul(y3, w3);
            }
            x[0] = y0; x[qs] = y1; x[2 * qs] = y2; x[3 * qs] = y3;
          } else {
            double2 v[13], y[13];
            for (int q = 0; q < r; q++) v[q] = x[q * qs];
            for (int p = 0; p < r; p++) {
              double2 acc = v[0];
              for (int q = 1; q < r; q++) {
                double2 w = tw[((p * q) % r) * tr];
                if (inverse) w.y = -w.y;
                acc = cadd(acc, cmul(v[q], w));
              }
              if (p && j) {
                double2 w = tw[(int)(((long long)p * j * tL) % n)];
                if (inverse) w.y = -w.y;
                acc = cmul(acc, w);
              }
              y[p] = acc;
            }
            for (int p = 0; p < r; p++) x[p * qs] = y[p];
          }
        }
      }
      __syncthreads();
      L = M;
    }
  }

  // ---- phase 3: split (forward) / unpack (inverse) + store ----
  if (!inverse) {
    // X[k] = E + T, X[h-k] = conj(E - T), E = (Zk + conj Zhk)/2, T = -i/2 w^k (Zk - conj Zhk)
    const int npair = h / 2 + 1;
    const double fac = pl.fac;
    const int total = CONTIG ? nb * npair : npair;
    for (int idx = CONTIG ? tid : ty; idx < total; idx += CONTIG ? nthr : FFT_TY) {
      int b, k;
      if (CONTIG) { b = idx / npair; k = idx - b * npair; } else { b = lane; k = idx; if (lane >= nb) break; }
      double *q = out + obase + b * dd.s1;
      if (k == 0) {
        const double2 Z0 = buf[b];
        q[0] = (Z0.x + Z0.y) * fac;
        q[(long long)(n - 1) * dd.sp] = (Z0.x - Z0.y) * fac;
      } else {
        const double2 Zk = buf[fft_perm(pl, k) * FFT_BP + b];
        const double2 Zc = cconj(buf[fft_perm(pl, h - k) * FFT_BP + b]);
        const double2 E = make_double2(0.5 * (Zk.x + Zc.x), 0.5 * (Zk.y + Zc.y));
        const double2 D = csub(Zk, Zc);
        const double2 w = tw[k];
        // -i/2 * w = (w.y/2, -w.x/2)
        const double2 T = cmul(make_double2(0.5 * w.y, -0.5 * w.x), D);
        const double2 Xk = cadd(E, T), Xh = cconj(csub(E, T));
        q[(2 * k - 1) * dd.sp] = Xk.x * fac;
        q[(2 * k) * dd.sp] = Xk.y * fac;
        q[(2 * (h - k) - 1) * dd.sp] = Xh.x * fac;
        q[(2 * (h - k)) * dd.sp] = Xh.y * fac;
      }
    }
  } else {
    const double fac = pl.fac;
    if (CONTIG) {
      for (int idx = tid; idx < nb * h; idx += nthr) {
        const int b = idx / h, m = idx - b * h;
        const double2 z = buf[fft_perm(pl, m) * FFT_BP + b];
        double *q = out + obase + b * dd.s1 + 2 * m * dd.sp;
        q[0] = z.x * fac;
        q[dd.sp] = z.y * fac;
      }
    } else if (lane < nb) {
      for (int m = ty; m < h; m += FFT_TY) {
        const double2 z = buf[fft_perm(pl, m) * FFT_BP + lane];
        double *q = out + obase + lane * dd.s1 + 2 * m * dd.sp;
        q[0] = z.x * fac;
        q[dd.sp] = z.y * fac;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// solmpj: src/modpois.f90:1107-1166 with bxyzrt formed on the fly (src/modpois.f90:196-220):
// diag = b(k) + xrt(iglob) + yrt(jglob), except the global mean mode at k = ktot which uses
// b_top_D.  One thread per (i,j) column; d is a (imax,jmax,ktot) scratch like the reference's.
__global__ void __launch_bounds__(128) k_solmpj(Geo g, double *__restrict__ x, double *__restrict__ d,
                                                const double *__restrict__ xrt, const double *__restrict__ yrt,
                                                const double *__restrict__ a, const double *__restrict__ b,
                                                const double *__restrict__ c, double b_top_D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // 0-based local
  const int j = blockIdx.y;
  if (i >= g.imax) return;
  const int K = g.ktot;
  const double lam = xrt[g.i0g + i] + yrt[g.j0g + j];
  const long long sk = (long long)g.imax * g.jmax;
  long long q = (long long)i + (long long)g.imax * j;
  double bb = (lam == 0. && K == 1) ? b_top_D : b[0] + lam;
  double z = 1. / bb;
  double dk = c[0] * z;
  double xk = x[q] * z;
  d[q] = dk; x[q] = xk;
  for (int k = 1; k < K - 1; k++) {
    q += sk;
    bb = b[k] + lam;
    z = 1. / (bb - a[k] * dk);
    dk = c[k] * z;
    xk = (x[q] - a[k] * xk) * z;
    d[q] = dk; x[q] = xk;
  }
  if (K > 1) {
    q += sk;
    bb = (lam == 0.) ? b_top_D : b[K - 1] + lam;
    z = bb - a[K - 1] * dk;
    xk = (x[q] - a[K - 1] * xk) / z;
    x[q] = xk;
  }
  for (int k = K - 2; k >= 0; k--) {
    q -= sk;
    xk = x[q] - d[q] * xk;
    x[q] = xk;
  }
}

}  // namespace udg
