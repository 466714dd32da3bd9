// common.cuh — shared definitions for the sm_100a uDALES dynamics kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace udg {

// Geometry + metric tables of one z-pencil, passed by value to every kernel.
// Storage indices are 0-based: Fortran (i,j,k) of a momentum-halo array lives at
// (i+ih-1) + pi*((j+jh-1) + pj*(k+kh-1))   (src/modfields.f90:440-442 shapes).
struct Geo {
  int imax, jmax, ktot;
  int itot, jtot;
  int ih, jh, kh;
  int pi, pj;            // pitches of momentum-halo arrays: imax+2ih, jmax+2jh
  long long pk;          // pi*pj
  int ihc, jhc, khc, pic, pjc;
  long long pkc;
  int i0g, j0g;          // global 0-based offset of local cell 1 (zstart-1)
  double dx, dy, dxi, dyi, dxiq, dyiq, dx2i, dy2i, dx2, dy2, dxi5, dyi5;
  // 1-D metric tables on the device, pre-offset so that [k] is valid for Fortran k in -1..ktot+2
  const double *dzf, *dzh, *dzfi, *dzhi, *dzhiq, *dzfiq, *dzh2i, *dzf2, *dzfi5, *delta;
  const double *dzfc, *dzfci, *dzhci;
  double numol, prandtlmoli, prandtli, c_vreman, csz;
  double Uinf, Vinf;
  int BCtopm, lles;
  int wrapx;             // 1: x is unsplit and kernels that own their halos write the periodic x images themselves
  int xalt;              // 1: x-spectral slots between the two x transforms are aligned pairs (poisson_xline.cuh), not the packed order
};

// offset of Fortran element (i,j,k) in a momentum-halo array starting at k = 1-kh
__host__ __device__ __forceinline__ long long offF(const Geo &g, int i, int j, int k) {
  return (long long)(i + g.ih - 1) + (long long)g.pi * ((j + g.jh - 1) + (long long)g.pj * (k + g.kh - 1));
}
// tendency-type array starting at k = 1
__host__ __device__ __forceinline__ long long offT(const Geo &g, int i, int j, int k) {
  return (long long)(i + g.ih - 1) + (long long)g.pi * ((j + g.jh - 1) + (long long)g.pj * (k - 1));
}
// halo-free (imax,jmax,ktot)
__host__ __device__ __forceinline__ long long offR(const Geo &g, int i, int j, int k) {
  return (long long)(i - 1) + (long long)g.imax * ((j - 1) + (long long)g.jmax * (k - 1));
}
__host__ __device__ __forceinline__ long long offS(const Geo &g, int i, int j, int k) {
  return (long long)(i + g.ihc - 1) + (long long)g.pic * ((j + g.jhc - 1) + (long long)g.pjc * (k + g.khc - 1));
}
// Periodic images of interior cell (i,j) in the halo ring (width 1, imax,jmax >= 2): at most one in x
// (only when x is unsplit), one in y (nprocy = 1 always) and the corner.  -1 = none.
__device__ __forceinline__ int img_x(const Geo &g, int i) { return g.wrapx ? (i == 1 ? g.imax + 1 : (i == g.imax ? 0 : -1)) : -1; }
__device__ __forceinline__ int img_y(const Geo &g, int j) { return j == 1 ? g.jmax + 1 : (j == g.jmax ? 0 : -1); }

__host__ __device__ __forceinline__ long long offST(const Geo &g, int i, int j, int k) {
  return (long long)(i + g.ihc - 1) + (long long)g.pic * ((j + g.jhc - 1) + (long long)g.pjc * (k - 1));
}

}  // namespace udg
