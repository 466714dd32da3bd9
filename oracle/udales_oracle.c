/* udales_oracle.c — CPU restatement (parity ORACLE) of the uDALES dynamics hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see udales_oracle.h).  Plain C + OpenMP, fp64 like the
 * reference (-fdefault-real-8, CMakeLists.txt:46).  Build with -ffp-contract=off so
 * that, like the reference's generic x86-64 -O3 build, no FMA contraction happens.
 *
 * PINNING STATUS: pinned to the reference's own source text.  tests/golden/ref_*.npz are produced by
 * oracle/f90run (an interpreter that executes the .f90 files under /root/reference/src on seeded inputs; see
 * oracle/README.md) and tests/test_oracle_golden.py checks every stage of three RK3 substeps of this
 * file against them (<= 2e-13 for the stencils, 1e-11 through the FFT solve).  The reference itself is
 * unbuildable here (no Fortran compiler / MPI / FFTW) and ships no golden vectors of its own; FFTW is
 * represented by numpy.fft in the golden generator.
 */
#include "udales_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
struct orc {
  orc_cfg c;
  int itot, jtot, ktot;       /* == imax, jmax, kmax (single pencil)          */
  int ih, jh, kh;             /* momentum halo: 1,1,1 (modglobal.f90:592-599) */
  int ihc, jhc, khc;          /* scalar halo: 2 with kappa (modglobal.f90:602-609) */
  int nsv;
  double dx, dy, dxi, dyi, dx2i, dy2i, dxiq, dyiq, dxi5, dyi5, dx2, dy2;
  /* 1-D metrics, stored with offset MOFF so index k = -1 .. ktot+2 is valid   */
  double *zf, *zh, *dzf, *dzh, *dzfi, *dzhi, *dzf2, *dzhiq, *dzfiq, *dzh2i, *dzfi5;
  double *dzfc, *dzfci, *dzhci; /* kappa metrics (modglobal.f90:841-870)       */
  double *delta;                /* delta(k) (x uniform)                        */
  double *xrt, *yrt, *a, *b, *cc; /* 1-based via -1 offset macros              */
  double b_top_D;
  /* 3-D fields */
  double *u0, *v0, *w0, *um, *vm, *wm, *pres0, *p, *ekm, *ekh; /* (1-ih:,1-jh:,1-kh:ktot+kh) */
  double *up, *vp, *wp, *pup, *pvp, *pwp;                        /* (.., .., 1:ktot+kh)        */
  double *rhs;                                                   /* (itot,jtot,ktot)           */
  double *sv0, *svm, *svp;    /* (1-ihc:,1-jhc:,1-khc:ktot+khc | 1:ktot+khc, nsv)               */
  double *d;                  /* solmpj scratch (itot,jtot,ktot)                                */
  double *dumu, *duml;        /* advecc_kappa temporaries                                       */
  /* FFT tables */
  double *twx, *twy;
  /* immersed boundary (src/modibm.f90): local 1-based point lists + real masks (1 fluid, 0 solid) */
  int ibm_n[8];               /* 0-3 solid_u,v,w,c ; 4-7 fluid-boundary points u,v,w,c */
  int *ibm_pts[8];            /* 3*n, point-major: (i,j,k) of point n at [3n..3n+2]    */
  double *mask[4];            /* mask_u, mask_v, mask_w, mask_c  (momentum-halo shape)  */
  int libm;
  /* forces (src/modforces.f90:46): large-scale pressure gradient per level, kb:ke+kh */
  double *dpdxl, *dpdyl;
  int has_forcing;
  /* bottom -> wfmneutral (src/modibm.f90:1998, src/modwallfunctions.f90:262) and masscorr (src/modforces.f90:328) */
  int lbottom, BCbotm, BCbots;
  double z0, fkar;
  double *momfluxb;
  int luvolflowr, lvvolflowr;
  double uflowrate, vflowrate, udef, vdef;
  int *IIu, *IIv;             /* (itot, jtot, ktot+1), 1 = fluid (createmasks, src/modibm.f90:2103); NULL = all fluid */
  int *IIus, *IIvs;           /* fluid points per level, ktot+1 values */
  /* temperature (ltempeq, dry: lmoist = .false.) */
  int ltempeq, lbuoyancy, BCtopT, BCbotT, lbuoycorr;
  double grav, thls, wttop, thl_top, wtsurf, Rigc;
  /* wfuno (BCbotm = 2 / BCbotT = 2): z0h, prandtlturb, gravity, wall temperature; tcell = the uniform thl0(kb) of a run
   * without temperature equation (thl0 stays at thlprof there) */
  double wf_z0h, wf_prandtlturb, wf_grav, wf_twall, wf_tcell;
  double *thl0, *thlm;        /* momentum-halo shape (alloc_z, src/modfields.f90:495-497) */
  double *thlp;               /* tendency shape (src/modfields.f90:453) */
  double *thl0h;              /* momentum-halo shape (:500); interior only is ever written */
  double *thv0h, *dthvdz;     /* tendency shape (:464-465) */
  double *thl0av, *thvh, *thlpcar; /* kb:ke+kh */
};

#define MOFF 2
/* momentum-halo arrays starting at k = 1-kh */
#define PI_ (o->itot + 2 * o->ih)
#define PJ_ (o->jtot + 2 * o->jh)
#define F(a, i, j, k) (a)[((i) + o->ih - 1) + (size_t)PI_ * (((j) + o->jh - 1) + (size_t)PJ_ * ((k) + o->kh - 1))]
/* tendency-type arrays starting at k = 1 */
#define T(a, i, j, k) (a)[((i) + o->ih - 1) + (size_t)PI_ * (((j) + o->jh - 1) + (size_t)PJ_ * ((k)-1))]
/* halo-free arrays */
#define R(a, i, j, k) (a)[((i)-1) + (size_t)o->itot * (((j)-1) + (size_t)o->jtot * ((k)-1))]
/* scalar-halo arrays */
#define PIC_ (o->itot + 2 * o->ihc)
#define PJC_ (o->jtot + 2 * o->jhc)
#define S(a, i, j, k) (a)[((i) + o->ihc - 1) + (size_t)PIC_ * (((j) + o->jhc - 1) + (size_t)PJC_ * ((k) + o->khc - 1))]
#define ST(a, i, j, k) (a)[((i) + o->ihc - 1) + (size_t)PIC_ * (((j) + o->jhc - 1) + (size_t)PJC_ * ((k)-1))]
#define M(a, k) (o->a)[(k) + MOFF]

static size_t nF(const orc_t *o) { return (size_t)(o->itot + 2 * o->ih) * (o->jtot + 2 * o->jh) * (o->ktot + 2 * o->kh); }
static size_t nT(const orc_t *o) { return (size_t)(o->itot + 2 * o->ih) * (o->jtot + 2 * o->jh) * (o->ktot + o->kh); }
static size_t nR(const orc_t *o) { return (size_t)o->itot * o->jtot * o->ktot; }
static size_t nS(const orc_t *o) { return (size_t)(o->itot + 2 * o->ihc) * (o->jtot + 2 * o->jhc) * (o->ktot + 2 * o->khc); }
static size_t nST(const orc_t *o) { return (size_t)(o->itot + 2 * o->ihc) * (o->jtot + 2 * o->jhc) * (o->ktot + o->khc); }

static double *zalloc(size_t n) {
  double *p = (double *)calloc(n ? n : 1, sizeof(double));
  if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
  return p;
}

/* ------------------------------------------------------------------------- */
/* initglobal metrics: src/modglobal.f90:708-838 ; initpois: src/modpois.f90:91-176 */
orc_t *orc_create(const orc_cfg *cfg) {
  orc_t *o = (orc_t *)calloc(1, sizeof(orc_t));
  o->c = *cfg;
  o->itot = cfg->itot; o->jtot = cfg->jtot; o->ktot = cfg->ktot;
  o->nsv = cfg->nsv;
  o->ih = o->jh = o->kh = 1;                         /* modglobal.f90:592-599 */
  o->ihc = o->jhc = o->khc = (o->nsv > 0 && cfg->iadv_sv == 7) ? 2 : 1; /* :602-609 */
  const int K = o->ktot;
  o->dx = cfg->xlen / (double)o->itot;               /* :710 */
  o->dy = cfg->ylen / (double)o->jtot;
  const int nm = K + 2 * MOFF + 2;
  o->zf = zalloc(nm); o->zh = zalloc(nm); o->dzf = zalloc(nm); o->dzh = zalloc(nm);
  o->dzfi = zalloc(nm); o->dzhi = zalloc(nm); o->dzf2 = zalloc(nm); o->dzhiq = zalloc(nm);
  o->dzfiq = zalloc(nm); o->dzh2i = zalloc(nm); o->dzfi5 = zalloc(nm);
  o->dzfc = zalloc(nm); o->dzfci = zalloc(nm); o->dzhci = zalloc(nm); o->delta = zalloc(nm);
  for (int k = 1; k <= K; k++) M(zf, k) = cfg->zf[k - 1];
  M(zh, 1) = 0.0;                                     /* :746 */
  for (int k = 1; k <= K; k++) M(zh, k + 1) = M(zh, k) + 2.0 * (M(zf, k) - M(zh, k));
  M(zf, K + 1) = M(zf, K) + 2.0 * (M(zh, K + 1) - M(zf, K));
  for (int k = 1; k <= K; k++) M(dzf, k) = M(zh, k + 1) - M(zh, k);
  M(dzf, K + 1) = M(dzf, K);
  M(dzf, 0) = M(dzf, 1);
  M(dzh, 1) = 2 * M(zf, 1);
  for (int k = 2; k <= K + 1; k++) M(dzh, k) = M(zf, k) - M(zf, k - 1);
  for (int k = 0; k <= K + 1; k++) {
    M(dzfi, k) = 1. / M(dzf, k);
    M(dzf2, k) = M(dzf, k) * M(dzf, k);
    M(dzfiq, k) = 0.25 * M(dzfi, k);
    M(dzfi5, k) = 0.5 * M(dzfi, k);
  }
  for (int k = 1; k <= K + 1; k++) {
    M(dzhi, k) = 1. / M(dzh, k);
    M(dzhiq, k) = 0.25 * M(dzhi, k);
    M(dzh2i, k) = M(dzhi, k) * M(dzhi, k);
    M(delta, k) = pow(o->dx * o->dy * M(dzf, k), 1. / 3.);  /* :793-797 (dxf == dx) */
  }
  o->dxi = 1. / o->dx; o->dyi = 1. / o->dy;
  o->dx2 = o->dx * o->dx; o->dy2 = o->dy * o->dy;
  o->dxiq = 0.25 * o->dxi; o->dyiq = 0.25 * o->dyi;
  o->dx2i = o->dxi * o->dxi; o->dy2i = o->dyi * o->dyi;
  o->dxi5 = 0.5 * o->dxi; o->dyi5 = 0.5 * o->dyi;
  /* kappa metrics, modglobal.f90:841-870: dzfc = dzf extended by one more ghost, dzhci likewise */
  for (int k = 0; k <= K + 1; k++) M(dzfc, k) = M(dzf, k);
  M(dzfc, -1) = M(dzfc, 0);
  M(dzfc, K + 2) = M(dzfc, K + 1);
  for (int k = -1; k <= K + 2; k++) M(dzfci, k) = 1. / M(dzfc, k);
  for (int k = 1; k <= K + 1; k++) M(dzhci, k) = M(dzhi, k);
  M(dzhci, 0) = M(dzhci, 1);
  M(dzhci, K + 2) = M(dzhci, K + 1);

  /* initpois FFT2D, periodic x/y, BCzp == 1 : src/modpois.f90:98-176 */
  const double pi = 3.141592653589793116;            /* modglobal.f90 pi     */
  o->xrt = zalloc(o->itot + 1); o->yrt = zalloc(o->jtot + 1);
  o->a = zalloc(K + 2); o->b = zalloc(K + 2); o->cc = zalloc(K + 2);
  {
    double fac = 1. / (2. * o->itot);
    for (int i = 3; i <= o->itot; i += 2) {
      double s = sin((double)(i - 1) * pi * fac);
      o->xrt[i - 1] = -4. * o->dxi * o->dxi * (s * s);
      o->xrt[i] = o->xrt[i - 1];
    }
    o->xrt[1] = 0.;
    o->xrt[o->itot] = -4. * o->dxi * o->dxi;
    fac = 1. / (2. * o->jtot);
    for (int j = 3; j <= o->jtot; j += 2) {
      double s = sin((double)(j - 1) * pi * fac);
      o->yrt[j - 1] = -4. * o->dyi * o->dyi * (s * s);
      o->yrt[j] = o->yrt[j - 1];
    }
    o->yrt[1] = 0.;
    o->yrt[o->jtot] = -4. * o->dyi * o->dyi;
  }
  for (int k = 1; k <= K; k++) {                       /* rhobf = rhobh = 1 (modfields.f90:571-572) */
    o->a[k] = 1. / (M(dzf, k) * M(dzh, k));
    o->cc[k] = 1. / (M(dzf, k) * M(dzh, k + 1));
    o->b[k] = -(o->a[k] + o->cc[k]);
  }
  o->b[1] = o->b[1] + o->a[1];
  {
    double b_top_N = o->b[K] + o->cc[K];
    o->b_top_D = o->b[K] - o->cc[K];
    o->b[K] = b_top_N;                                /* kbc2 == 1 */
  }
  o->a[1] = 0.;
  o->cc[K] = 0.;

  o->u0 = zalloc(nF(o)); o->v0 = zalloc(nF(o)); o->w0 = zalloc(nF(o));
  o->um = zalloc(nF(o)); o->vm = zalloc(nF(o)); o->wm = zalloc(nF(o));
  o->pres0 = zalloc(nF(o)); o->p = zalloc(nF(o)); o->ekm = zalloc(nF(o)); o->ekh = zalloc(nF(o));
  o->up = zalloc(nT(o)); o->vp = zalloc(nT(o)); o->wp = zalloc(nT(o));
  o->pup = zalloc(nT(o)); o->pvp = zalloc(nT(o)); o->pwp = zalloc(nT(o));
  o->rhs = zalloc(nR(o)); o->d = zalloc(nR(o));
  if (o->nsv > 0) {
    o->sv0 = zalloc(nS(o) * o->nsv); o->svm = zalloc(nS(o) * o->nsv); o->svp = zalloc(nST(o) * o->nsv);
    o->dumu = zalloc(nST(o)); o->duml = zalloc(nST(o));
  }
  return o;
}

void orc_destroy(orc_t *o) {
  if (!o) return;
  double *all[] = {o->zf, o->zh, o->dzf, o->dzh, o->dzfi, o->dzhi, o->dzf2, o->dzhiq, o->dzfiq, o->dzh2i, o->dzfi5,
                   o->dzfc, o->dzfci, o->dzhci, o->delta, o->xrt, o->yrt, o->a, o->b, o->cc,
                   o->u0, o->v0, o->w0, o->um, o->vm, o->wm, o->pres0, o->p, o->ekm, o->ekh,
                   o->up, o->vp, o->wp, o->pup, o->pvp, o->pwp, o->rhs, o->d, o->sv0, o->svm, o->svp,
                   o->dumu, o->duml, o->twx, o->twy, o->thl0, o->thlm, o->thlp, o->thl0h, o->thv0h, o->dthvdz,
                   o->thl0av, o->thvh, o->thlpcar};
  for (size_t n = 0; n < sizeof(all) / sizeof(all[0]); n++) free(all[n]);
  free(o);
}

double *orc_field(orc_t *o, const char *name, int dims[4]) {
  struct { const char *n; double *p; int kind; } tab[] = {
      {"u0", o->u0, 0}, {"v0", o->v0, 0}, {"w0", o->w0, 0}, {"um", o->um, 0}, {"vm", o->vm, 0}, {"wm", o->wm, 0},
      {"pres0", o->pres0, 0}, {"p", o->p, 0}, {"ekm", o->ekm, 0}, {"ekh", o->ekh, 0},
      {"up", o->up, 1}, {"vp", o->vp, 1}, {"wp", o->wp, 1}, {"pup", o->pup, 1}, {"pvp", o->pvp, 1}, {"pwp", o->pwp, 1},
      {"rhs", o->rhs, 2}, {"sv0", o->sv0, 3}, {"svm", o->svm, 3}, {"svp", o->svp, 4},
      {"thl0", o->thl0, 0}, {"thlm", o->thlm, 0}, {"thl0h", o->thl0h, 0}, {"thlp", o->thlp, 1}, {"thv0h", o->thv0h, 1},
      {"dthvdz", o->dthvdz, 1}};
  for (size_t n = 0; n < sizeof(tab) / sizeof(tab[0]); n++)
    if (!strcmp(tab[n].n, name)) {
      int kd = tab[n].kind;
      if (!tab[n].p) return NULL;
      dims[3] = 1;
      if (kd <= 1) { dims[0] = PI_; dims[1] = PJ_; dims[2] = o->ktot + (kd == 0 ? 2 : 1) * o->kh; }
      else if (kd == 2) { dims[0] = o->itot; dims[1] = o->jtot; dims[2] = o->ktot; }
      else { dims[0] = PIC_; dims[1] = PJC_; dims[2] = o->ktot + (kd == 3 ? 2 : 1) * o->khc; dims[3] = o->nsv; }
      return tab[n].p;
    }
  return NULL;
}

double *orc_metric(orc_t *o, const char *name, int *lo, int *n) {
  const int K = o->ktot;
  struct { const char *nm; double *p; int lo, n; } tab[] = {
      {"dzf", &M(dzf, 0), 0, K + 2}, {"dzh", &M(dzh, 1), 1, K + 1}, {"dzfi", &M(dzfi, 0), 0, K + 2},
      {"dzhi", &M(dzhi, 1), 1, K + 1}, {"zf", &M(zf, 1), 1, K + 1}, {"zh", &M(zh, 1), 1, K + 1},
      {"delta", &M(delta, 1), 1, K + 1},
      {"xrt", o->xrt + 1, 1, o->itot}, {"yrt", o->yrt + 1, 1, o->jtot},
      {"a", o->a + 1, 1, K}, {"b", o->b + 1, 1, K}, {"c", o->cc + 1, 1, K}, {"b_top_D", &o->b_top_D, 1, 1}};
  for (size_t q = 0; q < sizeof(tab) / sizeof(tab[0]); q++)
    if (!strcmp(tab[q].nm, name)) { *lo = tab[q].lo; *n = tab[q].n; return tab[q].p; }
  return NULL;
}

/* ------------------------------------------------------------------------- */
/* advecu_2nd: src/modadvection.f90:158-212 */
static void advecu_2nd(orc_t *o, const double *putin, double *putout) {
  const double *u0 = o->u0, *v0 = o->v0, *w0 = o->w0, *pres0 = o->pres0;
  const double dxi = o->dxi, dxiq = o->dxiq, dyiq = o->dyiq;
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= o->ktot; k++)
    for (int j = 1; j <= o->jtot; j++) {
      const int jm = j - 1, jp = j + 1;
      for (int i = 1; i <= o->itot; i++) {
        const int im = i - 1, ip = i + 1;
        T(putout, i, j, k) = T(putout, i, j, k) - (
            ((F(putin, i, j, k) + F(putin, ip, j, k)) * (F(u0, i, j, k) + F(u0, ip, j, k))
           - (F(putin, i, j, k) + F(putin, im, j, k)) * (F(u0, i, j, k) + F(u0, im, j, k))) * dxiq
          + ((F(putin, i, j, k) + F(putin, i, jp, k)) * (F(v0, i, jp, k) + F(v0, im, jp, k))
           - (F(putin, i, j, k) + F(putin, i, jm, k)) * (F(v0, i, j, k) + F(v0, im, j, k))) * dyiq)
          - ((F(pres0, i, j, k) - F(pres0, i - 1, j, k)) * dxi);
      }
    }
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= o->ktot; k++) {        /* reference is k-innermost (:193-210); per-cell independent */
    const int km = k - 1, kp = k + 1;
    for (int j = 1; j <= o->jtot; j++)
      for (int i = 1; i <= o->itot; i++) {
        const int im = i - 1;
        T(putout, i, j, k) = T(putout, i, j, k) - (
            (F(putin, i, j, kp) * M(dzf, k) + F(putin, i, j, k) * M(dzf, kp)) * M(dzhi, kp)
              * (F(w0, i, j, kp) + F(w0, im, j, kp))
          - (F(putin, i, j, k) * M(dzf, km) + F(putin, i, j, km) * M(dzf, k)) * M(dzhi, k)
              * (F(w0, i, j, k) + F(w0, im, j, k))) * 0.5 * M(dzfi5, k);
      }
  }
}

/* advecv_2nd: src/modadvection.f90:215-270 */
static void advecv_2nd(orc_t *o, const double *putin, double *putout) {
  const double *u0 = o->u0, *v0 = o->v0, *w0 = o->w0, *pres0 = o->pres0;
  const double dyi = o->dyi, dxiq = o->dxiq, dyiq = o->dyiq;
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= o->ktot; k++)
    for (int j = 1; j <= o->jtot; j++) {
      const int jm = j - 1, jp = j + 1;
      for (int i = 1; i <= o->itot; i++) {
        const int im = i - 1, ip = i + 1;
        T(putout, i, j, k) = T(putout, i, j, k) - (
            ((F(u0, ip, j, k) + F(u0, ip, jm, k)) * (F(putin, i, j, k) + F(putin, ip, j, k))
           - (F(u0, i, j, k) + F(u0, i, jm, k)) * (F(putin, i, j, k) + F(putin, im, j, k))) * dxiq
          + ((F(v0, i, jp, k) + F(v0, i, j, k)) * (F(putin, i, j, k) + F(putin, i, jp, k))
           - (F(v0, i, jm, k) + F(v0, i, j, k)) * (F(putin, i, j, k) + F(putin, i, jm, k))) * dyiq)
          - ((F(pres0, i, j, k) - F(pres0, i, jm, k)) * dyi);
      }
    }
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= o->ktot; k++) {
    const int km = k - 1, kp = k + 1;
    for (int j = 1; j <= o->jtot; j++) {
      const int jm = j - 1;
      for (int i = 1; i <= o->itot; i++) {
        T(putout, i, j, k) = T(putout, i, j, k) - (
            (F(w0, i, j, kp) + F(w0, i, jm, kp))
              * (F(putin, i, j, kp) * M(dzf, k) + F(putin, i, j, k) * M(dzf, kp)) * M(dzhi, kp)
          - (F(w0, i, j, k) + F(w0, i, jm, k))
              * (F(putin, i, j, km) * M(dzf, k) + F(putin, i, j, k) * M(dzf, km)) * M(dzhi, k)) * 0.5 * M(dzfi5, k);
      }
    }
  }
}

/* advecw_2nd: src/modadvection.f90:273-314 */
static void advecw_2nd(orc_t *o, const double *putin, double *putout) {
  const double *u0 = o->u0, *v0 = o->v0, *w0 = o->w0, *pres0 = o->pres0;
  const double dxiq = o->dxiq, dyiq = o->dyiq;
#pragma omp parallel for schedule(static)
  for (int k = 2; k <= o->ktot; k++) {
    const int km = k - 1, kp = k + 1;
    for (int j = 1; j <= o->jtot; j++) {
      const int jm = j - 1, jp = j + 1;
      for (int i = 1; i <= o->itot; i++) {
        const int im = i - 1, ip = i + 1;
        T(putout, i, j, k) = T(putout, i, j, k) - (
            ((F(putin, ip, j, k) + F(putin, i, j, k)) * (M(dzf, km) * F(u0, ip, j, k) + M(dzf, k) * F(u0, ip, j, km))
           - (F(putin, i, j, k) + F(putin, im, j, k)) * (M(dzf, km) * F(u0, i, j, k) + M(dzf, k) * F(u0, i, j, km))
            ) * dxiq * M(dzhi, k)
          + ((F(putin, i, jp, k) + F(putin, i, j, k)) * (M(dzf, km) * F(v0, i, jp, k) + M(dzf, k) * F(v0, i, jp, km))
           - (F(putin, i, j, k) + F(putin, i, jm, k)) * (M(dzf, km) * F(v0, i, j, k) + M(dzf, k) * F(v0, i, j, km))
            ) * dyiq * M(dzhi, k)
          + ((F(putin, i, j, k) + F(putin, i, j, kp)) * (F(w0, i, j, k) + F(w0, i, j, kp))
           - (F(putin, i, j, k) + F(putin, i, j, km)) * (F(w0, i, j, k) + F(w0, i, j, km))
            ) * M(dzhiq, k))
          - ((F(pres0, i, j, k) - F(pres0, i, j, km)) * M(dzhi, k));
      }
    }
  }
}

/* rlim: src/modadvection.f90:408-421 */
static inline double rlim(double d1, double d2) {
  const double eps1 = 1.e-10;                          /* modglobal.f90 eps1 */
  double ri = (d2 + eps1) / (d1 + eps1);
  double phir = fmax(0., fmin(2. * ri, fmin(1. / 3. + 2. / 3. * ri, 2.)));
  return 0.5 * phir * d1;
}

/* advecc_kappa: src/modadvection.f90:316-406 (x uniform: dxhci = dxi, dxfc = dx, dxfci = dxi) */
static void advecc_kappa(orc_t *o, const double *var, double *varp) {
  const double *u0 = o->u0, *v0 = o->v0, *w0 = o->w0;
  double *dumu = o->dumu, *duml = o->duml;
  const size_t n = nST(o);
  const double dxhci = o->dxi, dxfc = o->dx, dxfci = o->dxi, dyi = o->dyi;
  memset(dumu, 0, n * sizeof(double)); memset(duml, 0, n * sizeof(double));
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= o->ktot; k++)
    for (int j = 1; j <= o->jtot; j++)
      for (int i = 1; i <= o->itot + 1; i++) {
        double d1, d2, cf;
        if (F(u0, i, j, k) > 0) {
          d1 = (S(var, i - 1, j, k) - S(var, i - 2, j, k)) * dxhci;
          d2 = (S(var, i, j, k) - S(var, i - 1, j, k)) * dxhci;
          cf = S(var, i - 1, j, k);
        } else {
          d1 = (S(var, i, j, k) - S(var, i + 1, j, k)) * dxhci;
          d2 = (S(var, i - 1, j, k) - S(var, i, j, k)) * dxhci;
          cf = S(var, i, j, k);
        }
        cf = cf + dxfc * rlim(d1, d2);
        ST(dumu, i - 1, j, k) = -cf * F(u0, i, j, k) * dxfci;
        ST(duml, i, j, k) = cf * F(u0, i, j, k) * dxfci;
      }
#pragma omp parallel for schedule(static)
  for (size_t q = 0; q < n; q++) varp[q] = varp[q] + dumu[q] + duml[q];
  memset(dumu, 0, n * sizeof(double)); memset(duml, 0, n * sizeof(double));
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= o->ktot; k++)
    for (int j = 1; j <= o->jtot + 1; j++)
      for (int i = 1; i <= o->itot; i++) {
        double d1, d2, cf;
        if (F(v0, i, j, k) > 0) {
          d1 = S(var, i, j - 1, k) - S(var, i, j - 2, k);
          d2 = S(var, i, j, k) - S(var, i, j - 1, k);
          cf = S(var, i, j - 1, k);
        } else {
          d1 = S(var, i, j, k) - S(var, i, j + 1, k);
          d2 = S(var, i, j - 1, k) - S(var, i, j, k);
          cf = S(var, i, j, k);
        }
        cf = cf + rlim(d1, d2);
        ST(duml, i, j, k) = cf * F(v0, i, j, k) * dyi;
        ST(dumu, i, j - 1, k) = -cf * F(v0, i, j, k) * dyi;
      }
#pragma omp parallel for schedule(static)
  for (size_t q = 0; q < n; q++) varp[q] = varp[q] + dumu[q] + duml[q];
  memset(dumu, 0, n * sizeof(double)); memset(duml, 0, n * sizeof(double));
  /* k loop writes dumu(k-1) and duml(k): different k never collide within one array */
#pragma omp parallel for schedule(static)
  for (int k = 2; k <= o->ktot + 1; k++)
    for (int j = 1; j <= o->jtot; j++)
      for (int i = 1; i <= o->itot; i++) {
        double d1, d2, cf;
        if (F(w0, i, j, k) > 0) {
          d1 = (S(var, i, j, k - 1) - S(var, i, j, k - 2)) * M(dzhci, k - 1);
          d2 = (S(var, i, j, k) - S(var, i, j, k - 1)) * M(dzhci, k);
          cf = S(var, i, j, k - 1);
        } else {
          d1 = (S(var, i, j, k) - S(var, i, j, k + 1)) * M(dzhci, k + 1);
          d2 = (S(var, i, j, k - 1) - S(var, i, j, k)) * M(dzhci, k);
          cf = S(var, i, j, k);
        }
        cf = cf + M(dzfc, k) * rlim(d1, d2);
        ST(duml, i, j, k) = cf * F(w0, i, j, k) * M(dzfci, k);
        ST(dumu, i, j, k - 1) = -cf * F(w0, i, j, k) * M(dzfci, k - 1);
      }
#pragma omp parallel for schedule(static)
  for (size_t q = 0; q < n; q++) varp[q] = varp[q] + dumu[q] + duml[q];
}

/* advecc_2nd: src/modadvection.f90:103-155, on scalar-halo arrays */
static void advecc_2nd(orc_t *o, const double *putin, double *putout) {
  const double *u0 = o->u0, *v0 = o->v0, *w0 = o->w0;
  const double dxi5 = o->dxi5, dyi5 = o->dyi5;
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= o->ktot; k++) {
    const int km = k - 1, kp = k + 1;
    for (int j = 1; j <= o->jtot; j++) {
      const int jm = j - 1, jp = j + 1;
      for (int i = 1; i <= o->itot; i++) {
        const int im = i - 1, ip = i + 1;
        ST(putout, i, j, k) = ST(putout, i, j, k) - (
            (F(u0, ip, j, k) * (S(putin, ip, j, k) + S(putin, i, j, k))
           - F(u0, i, j, k) * (S(putin, im, j, k) + S(putin, i, j, k))) * dxi5
          + (F(v0, i, jp, k) * (S(putin, i, jp, k) + S(putin, i, j, k))
           - F(v0, i, j, k) * (S(putin, i, jm, k) + S(putin, i, j, k))) * dyi5);
        ST(putout, i, j, k) = ST(putout, i, j, k) - (
            F(w0, i, j, kp) * (S(putin, i, j, kp) * M(dzf, k) + S(putin, i, j, k) * M(dzf, kp)) * M(dzhi, kp)
          - F(w0, i, j, k) * (S(putin, i, j, km) * M(dzf, k) + S(putin, i, j, k) * M(dzf, km)) * M(dzhi, k)
            ) * M(dzfi5, k);
      }
    }
  }
}

/* advection: src/modadvection.f90:36-101 (iadv_mom = cd2; scalars kappa or cd2) */
void orc_advection(orc_t *o) {
  advecu_2nd(o, o->u0, o->up);
  advecv_2nd(o, o->v0, o->vp);
  advecw_2nd(o, o->w0, o->wp);
  for (int n = 0; n < o->nsv; n++) {
    if (o->c.iadv_sv == 7) advecc_kappa(o, o->sv0 + n * nS(o), o->svp + n * nST(o));
    else advecc_2nd(o, o->sv0 + n * nS(o), o->svp + n * nST(o));
  }
  if (o->ltempeq) {      /* iadv_thl = cd2: advecc_2nd(ih, jh, kh, thl0, thlp), src/modadvection.f90:67-69 */
    const int hc = o->ihc;
    o->ihc = o->jhc = o->khc = 1;       /* thl carries the momentum halo: S / ST index like F / T */
    advecc_2nd(o, o->thl0, o->thlp);
    o->ihc = o->jhc = o->khc = hc;
  }
}

/* ------------------------------------------------------------------------- */
/* fluxtop with flux = 0: src/modboundary.f90:1494-1505 ; valuetop :1507-1517 */
static void fluxtop0(orc_t *o, double *f) {
  const int K = o->ktot;
  for (int j = 1 - o->jh; j <= o->jtot + o->jh; j++)
    for (int i = 1 - o->ih; i <= o->itot + o->ih; i++) F(f, i, j, K + 1) = F(f, i, j, K);
}
/* fluxtop with any flux: src/modboundary.f90:1494-1508 */
static void fluxtop(orc_t *o, double *f, const double *ek, double flux) {
  const int K = o->ktot;
  if (fabs(flux) <= 1.e-10) { fluxtop0(o, f); return; }
  for (int j = 1 - o->jh; j <= o->jtot + o->jh; j++)
    for (int i = 1 - o->ih; i <= o->itot + o->ih; i++)
      F(f, i, j, K + 1) = F(f, i, j, K) + M(dzh, K + 1) * flux / (M(dzhi, K + 1) * (0.5 * (M(dzf, K) * F(ek, i, j, K + 1) + M(dzf, K + 1) * F(ek, i, j, K))));
}
static void valuetop(orc_t *o, double *f, double val) {
  const int K = o->ktot;
  for (int j = 1 - o->jh; j <= o->jtot + o->jh; j++)
    for (int i = 1 - o->ih; i <= o->itot + o->ih; i++) F(f, i, j, K + 1) = 2 * val - F(f, i, j, K);
}

/* closurebc: src/modboundary.f90:434-505, single pencil, periodic x and y */
static void closurebc(orc_t *o) {
  double *ekm = o->ekm, *ekh = o->ekh;
  const int I = o->itot, J = o->jtot, K = o->ktot;
  const double numol = o->c.numol, prandtlmoli = o->c.prandtlmoli;
  /* exchange_halo_z on one pencil with non-periodic communicators is a no-op */
  if (o->c.BCtopm == 1 || o->c.BCtopm == 3) {
    for (int j = 0; j <= J + 1; j++)
      for (int i = 0; i <= I + 1; i++) {
        F(ekm, i, j, K + 1) = F(ekm, i, j, K);
        F(ekh, i, j, K + 1) = F(ekh, i, j, K);
        F(ekm, i, j, 0) = 2. * numol - F(ekm, i, j, 1);
        F(ekh, i, j, 0) = (2. * numol * prandtlmoli) - F(ekh, i, j, 1);
      }
  } else if (o->c.BCtopm == 2) {
    for (int j = 0; j <= J + 1; j++)
      for (int i = 0; i <= I + 1; i++) {
        F(ekm, i, j, K + 1) = 2. * numol - F(ekm, i, j, K);
        F(ekh, i, j, K + 1) = (2. * numol * prandtlmoli) - F(ekh, i, j, K);
        F(ekm, i, j, 0) = 2. * numol - F(ekm, i, j, 1);
        F(ekh, i, j, 0) = (2. * numol * prandtlmoli) - F(ekh, i, j, 1);
      }
  }
  for (int k = 0; k <= K + 1; k++)                    /* :476-481 */
    for (int j = 0; j <= J + 1; j++) {
      F(ekm, 0, j, k) = F(ekm, I, j, k);
      F(ekm, I + 1, j, k) = F(ekm, 1, j, k);
      F(ekh, 0, j, k) = F(ekh, I, j, k);
      F(ekh, I + 1, j, k) = F(ekh, 1, j, k);
    }
  for (int k = 0; k <= K + 1; k++)                    /* :494-500 */
    for (int i = 0; i <= I + 1; i++) {
      F(ekm, i, 0, k) = F(ekm, i, J, k);
      F(ekm, i, J + 1, k) = F(ekm, i, 1, k);
      F(ekh, i, 0, k) = F(ekh, i, J, k);
      F(ekh, i, J + 1, k) = F(ekh, i, 1, k);
    }
  /* reassure_fluxtop_boundary: modboundary.f90:392-431 */
  if (o->c.BCtopm == 1 || o->c.BCtopm == 3) {
    fluxtop0(o, o->um); fluxtop0(o, o->u0); fluxtop0(o, o->vm); fluxtop0(o, o->v0);
  }
  if (o->ltempeq && o->BCtopT == 1) { fluxtop(o, o->thlm, o->ekh, o->wttop); fluxtop(o, o->thl0, o->ekh, o->wttop); }   /* :417-420 */
  /* nsv > 0 with BCtops flux (default, wsvtop = 0): fluxtopscal, see orc_boundary */
}

static inline double sq_(double x) { return x * x; }

/* closure: src/modsubgrid.f90:159-412 (Smagorinsky :208-267, Vreman :269-360, DNS :401-404) */
void orc_closure(orc_t *o) {
  const double *u0 = o->u0, *v0 = o->v0, *w0 = o->w0;
  double *ekm = o->ekm, *ekh = o->ekh;
  const double dxi = o->dxi, dyi = o->dyi, dxiq = o->dxiq, dyiq = o->dyiq, dx2 = o->dx2, dy2 = o->dy2;
  const double numol = o->c.numol, prandtlmoli = o->c.prandtlmoli, prandtli = o->c.prandtli;
  const size_t n = nF(o);
  if (o->c.lsmagorinsky) {
    /* csz: modsubgrid.f90:65-77 */
    const double pi = 3.141592653589793116, cf = 2.5, alpha_kolm = 1.5;
    const double cm = cf / (2. * pi) * pow(1.5 * alpha_kolm, -1.5);
    const double ceps = 2. * pi / cf * pow(1.5 * alpha_kolm, -1.5);
    const double csz = (o->c.cs == -1.) ? pow(cm * cm * cm / ceps, 0.25) : o->c.cs;
#pragma omp parallel for schedule(static)
    for (int k = 1; k <= o->ktot; k++) {
      const int kp = k + 1, km = k - 1;
      const double mlen = csz * M(delta, k);
      for (int j = 1; j <= o->jtot; j++) {
        const int jp = j + 1, jm = j - 1;
        for (int i = 1; i <= o->itot; i++) {
          const int ip = i + 1, im = i - 1;
          const double damp = 1.;
          double strain2;
#define SQ(x) sq_(x)
          strain2 = SQ((F(u0, ip, j, k) - F(u0, i, j, k)) * dxi)
                  + SQ((F(v0, i, jp, k) - F(v0, i, j, k)) * dyi)
                  + SQ((F(w0, i, j, kp) - F(w0, i, j, k)) * M(dzfi, k));
          strain2 = strain2 + 0.125 * (
              + SQ((F(w0, i, j, kp) - F(w0, im, j, kp)) * dxi + (F(u0, i, j, kp) - F(u0, i, j, k)) * M(dzhi, kp))
              + SQ((F(w0, i, j, k) - F(w0, im, j, k)) * dxi + (F(u0, i, j, k) - F(u0, i, j, km)) * M(dzhi, k))
              + SQ((F(w0, ip, j, k) - F(w0, i, j, k)) * dxi + (F(u0, ip, j, k) - F(u0, ip, j, km)) * M(dzhi, k))
              + SQ((F(w0, ip, j, kp) - F(w0, i, j, kp)) * dxi + (F(u0, ip, j, kp) - F(u0, ip, j, k)) * M(dzhi, kp)));
          strain2 = strain2 + 0.125 * (
              + SQ((F(u0, i, jp, k) - F(u0, i, j, k)) * dyi + (F(v0, i, jp, k) - F(v0, im, jp, k)) * dxi)
              + SQ((F(u0, i, j, k) - F(u0, i, jm, k)) * dyi + (F(v0, i, j, k) - F(v0, im, j, k)) * dxi)
              + SQ((F(u0, ip, j, k) - F(u0, ip, jm, k)) * dyi + (F(v0, ip, j, k) - F(v0, i, j, k)) * dxi)
              + SQ((F(u0, ip, jp, k) - F(u0, ip, j, k)) * dyi + (F(v0, ip, jp, k) - F(v0, i, jp, k)) * dxi));
          strain2 = strain2 + 0.125 * (
              + SQ((F(v0, i, j, kp) - F(v0, i, j, k)) * M(dzhi, kp) + (F(w0, i, j, kp) - F(w0, i, jm, kp)) * dyi)
              + SQ((F(v0, i, j, k) - F(v0, i, j, km)) * M(dzhi, k) + (F(w0, i, j, k) - F(w0, i, jm, k)) * dyi)
              + SQ((F(v0, i, jp, k) - F(v0, i, jp, km)) * M(dzhi, k) + (F(w0, i, jp, k) - F(w0, i, j, k)) * dyi)
              + SQ((F(v0, i, jp, kp) - F(v0, i, jp, k)) * M(dzhi, kp) + (F(w0, i, jp, kp) - F(w0, i, j, kp)) * dyi));
#undef SQ
          F(ekm, i, j, k) = ((mlen * damp) * (mlen * damp)) * sqrt(2. * strain2);
          F(ekh, i, j, k) = F(ekm, i, j, k) * prandtli;
        }
      }
    }
    for (size_t q = 0; q < n; q++) { ekm[q] = ekm[q] + numol; ekh[q] = ekh[q] + numol * prandtlmoli; }
  } else if (o->c.lvreman) {
#pragma omp parallel for schedule(static)
    for (int k = 1; k <= o->ktot; k++) {
      const int kp = k + 1, km = k - 1;
      for (int j = 1; j <= o->jtot; j++) {
        const int jp = j + 1, jm = j - 1;
        for (int i = 1; i <= o->itot; i++) {
          const int ip = i + 1, im = i - 1;
          double a11 = (F(u0, ip, j, k) - F(u0, i, j, k)) * dxi;
          double a12 = (F(v0, ip, jp, k) + F(v0, ip, j, k) - F(v0, im, jp, k) - F(v0, im, j, k)) * dxiq;
          double a13 = (F(w0, ip, j, kp) + F(w0, ip, j, k) - F(w0, im, j, kp) - F(w0, im, j, k)) * dxiq;
          double a21 = (F(u0, ip, jp, k) + F(u0, i, jp, k) - F(u0, ip, jm, k) - F(u0, i, jm, k)) * dyiq;
          double a22 = (F(v0, i, jp, k) - F(v0, i, j, k)) * dyi;
          double a23 = (F(w0, i, jp, kp) + F(w0, i, jp, k) - F(w0, i, jm, kp) - F(w0, i, jm, k)) * dyiq;
          double a31 = (
              ((F(u0, ip, j, kp) + F(u0, i, j, kp)) * M(dzf, k) + (F(u0, ip, j, k) + F(u0, i, j, k)) * M(dzf, kp)) * M(dzhi, kp)
            - ((F(u0, ip, j, k) + F(u0, i, j, k)) * M(dzf, km) + (F(u0, ip, j, km) + F(u0, i, j, km)) * M(dzf, k)) * M(dzhi, k))
            * M(dzfiq, k);
          double a32 = (
              ((F(v0, i, jp, kp) + F(v0, i, j, kp)) * M(dzf, k) + (F(v0, i, jp, k) + F(v0, i, j, k)) * M(dzf, kp)) * M(dzhi, kp)
            - ((F(v0, i, jp, k) + F(v0, i, j, k)) * M(dzf, km) + (F(v0, i, jp, km) + F(v0, i, j, km)) * M(dzf, k)) * M(dzhi, k))
            * M(dzfiq, k);
          double a33 = (F(w0, i, j, kp) - F(w0, i, j, k)) * M(dzfi, k);
          double aa = a11 * a11 + a21 * a21 + a31 * a31 + a12 * a12 + a22 * a22 + a32 * a32 + a13 * a13 + a23 * a23 + a33 * a33;
          const double dzf2 = M(dzf2, k);
          double b11 = dx2 * a11 * a11 + dy2 * a21 * a21 + dzf2 * a31 * a31;
          double b22 = dx2 * a12 * a12 + dy2 * a22 * a22 + dzf2 * a32 * a32;
          double b12 = dx2 * a11 * a12 + dy2 * a21 * a22 + dzf2 * a31 * a32;
          double b33 = dx2 * a13 * a13 + dy2 * a23 * a23 + dzf2 * a33 * a33;
          double b13 = dx2 * a11 * a13 + dy2 * a21 * a23 + dzf2 * a31 * a33;
          double b23 = dx2 * a12 * a13 + dy2 * a22 * a23 + dzf2 * a32 * a33;
          double bb = b11 * b22 - b12 * b12 + b11 * b33 - b13 * b13 + b22 * b33 - b23 * b23;
          if (bb < 1.e-8) F(ekm, i, j, k) = 0.;
          else F(ekm, i, j, k) = o->c.c_vreman * sqrt(bb / aa);
        }
      }
    }
    if (o->lbuoyancy && o->lbuoycorr) {   /* buoyancy correction of the Vreman model for stable stratification, :332-354 */
      const double *thl0 = o->thl0, *dthvdz = o->dthvdz;
      for (int k = 1; k <= o->ktot; k++) {
        const int kp = k + 1, km = k - 1;
        for (int j = 1; j <= o->jtot; j++) {
          const int jp = j + 1;
          for (int i = 1; i <= o->itot; i++) {
            const int ip = i + 1;
            const double du0dz = 0.5 * ((F(u0, i, j, kp) + F(u0, ip, j, kp)) - (F(u0, i, j, km) + F(u0, ip, j, km))) / (M(dzh, kp) + M(dzh, k));
            const double dv0dz = 0.5 * ((F(v0, i, j, kp) + F(v0, i, jp, kp)) - (F(v0, i, j, km) + F(v0, i, jp, km))) / (M(dzh, kp) + M(dzh, k));
            const double Rig = ((o->grav / F(thl0, i, j, k)) * T(dthvdz, i, j, k)) / (du0dz * du0dz + dv0dz * dv0dz + 1.e-10);
            F(ekm, i, j, k) = F(ekm, i, j, k) * sqrt(1.0 - fmin(fmax(Rig, 0.0), o->Rigc) / o->Rigc);
          }
        }
      }
    }
    for (size_t q = 0; q < n; q++) ekh[q] = ekm[q] * prandtli;
    for (size_t q = 0; q < n; q++) ekm[q] = ekm[q] + numol;
    for (size_t q = 0; q < n; q++) ekh[q] = ekh[q] + numol * prandtlmoli;
  } else {
    for (size_t q = 0; q < n; q++) { ekm[q] = numol; ekh[q] = numol * prandtlmoli; }
  }
  closurebc(o);
}

/* diffu: src/modsubgrid.f90:672-775 */
static void diffu(orc_t *o, double *putout) {
  const double *u0 = o->u0, *v0 = o->v0, *w0 = o->w0, *ekm = o->ekm;
  const double dxi = o->dxi, dyi = o->dyi, dx2i = o->dx2i, numol = o->c.numol;
  const int lles = o->c.lles;
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= o->ktot; k++) {
    const int kp = k + 1, km = k - 1;
    for (int j = 1; j <= o->jtot; j++) {
      const int jp = j + 1, jm = j - 1;
      for (int i = 1; i <= o->itot; i++) {
        if (lles) {
          double emom = (M(dzf, km) * (F(ekm, i, j, k) + F(ekm, i - 1, j, k)) +
                         M(dzf, k) * (F(ekm, i, j, km) + F(ekm, i - 1, j, km))) * M(dzhiq, k);
          double emop = (M(dzf, kp) * (F(ekm, i, j, k) + F(ekm, i - 1, j, k)) +
                         M(dzf, k) * (F(ekm, i, j, kp) + F(ekm, i - 1, j, kp))) * M(dzhiq, kp);
          double empo = 0.25 * ((F(ekm, i, j, k) + F(ekm, i, jp, k)) + (F(ekm, i - 1, j, k) + F(ekm, i - 1, jp, k)));
          double emmo = 0.25 * ((F(ekm, i, j, k) + F(ekm, i, jm, k)) + (F(ekm, i - 1, jm, k) + F(ekm, i - 1, j, k)));
          T(putout, i, j, k) = T(putout, i, j, k)
            + (F(ekm, i, j, k) * (F(u0, i + 1, j, k) - F(u0, i, j, k))
             - F(ekm, i - 1, j, k) * (F(u0, i, j, k) - F(u0, i - 1, j, k))) * 2. * dx2i
            + (empo * ((F(u0, i, jp, k) - F(u0, i, j, k)) * dyi + (F(v0, i, jp, k) - F(v0, i - 1, jp, k)) * dxi)
             - emmo * ((F(u0, i, j, k) - F(u0, i, jm, k)) * dyi + (F(v0, i, j, k) - F(v0, i - 1, j, k)) * dxi)) * dyi
            + (emop * ((F(u0, i, j, kp) - F(u0, i, j, k)) * M(dzhi, kp) + (F(w0, i, j, kp) - F(w0, i - 1, j, kp)) * dxi)
             - emom * ((F(u0, i, j, k) - F(u0, i, j, km)) * M(dzhi, k) + (F(w0, i, j, k) - F(w0, i - 1, j, k)) * dxi)) * M(dzfi, k);
        } else {
          T(putout, i, j, k) = T(putout, i, j, k)
            + (numol * (F(u0, i + 1, j, k) - F(u0, i, j, k)) * dxi
             - numol * (F(u0, i, j, k) - F(u0, i - 1, j, k)) * dxi) * 2. * dxi
            + (numol * ((F(u0, i, jp, k) - F(u0, i, j, k)) * dyi + (F(v0, i, jp, k) - F(v0, i - 1, jp, k)) * dxi)
             - numol * ((F(u0, i, j, k) - F(u0, i, jm, k)) * dyi + (F(v0, i, j, k) - F(v0, i - 1, j, k)) * dxi)) * dyi
            + (numol * ((F(u0, i, j, kp) - F(u0, i, j, k)) * M(dzhi, kp) + (F(w0, i, j, kp) - F(w0, i - 1, j, kp)) * dxi)
             - numol * ((F(u0, i, j, k) - F(u0, i, j, km)) * M(dzhi, k) + (F(w0, i, j, k) - F(w0, i - 1, j, k)) * dxi)) * M(dzfi, k);
        }
      }
    }
  }
}

/* diffv: src/modsubgrid.f90:778-886 */
static void diffv(orc_t *o, double *putout) {
  const double *u0 = o->u0, *v0 = o->v0, *w0 = o->w0, *ekm = o->ekm;
  const double dxi = o->dxi, dyi = o->dyi, dy2i = o->dy2i, numol = o->c.numol;
  const int lles = o->c.lles;
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= o->ktot; k++) {
    const int kp = k + 1, km = k - 1;
    for (int j = 1; j <= o->jtot; j++) {
      const int jp = j + 1, jm = j - 1;
      for (int i = 1; i <= o->itot; i++) {
        if (lles) {
          double eomm = (M(dzf, km) * (F(ekm, i, j, k) + F(ekm, i, jm, k)) +
                         M(dzf, k) * (F(ekm, i, j, km) + F(ekm, i, jm, km))) * M(dzhiq, k);
          double eomp = (M(dzf, kp) * (F(ekm, i, j, k) + F(ekm, i, jm, k)) +
                         M(dzf, k) * (F(ekm, i, j, kp) + F(ekm, i, jm, kp))) * M(dzhiq, kp);
          double emmo = 0.25 * (F(ekm, i, j, k) + F(ekm, i, jm, k) + F(ekm, i - 1, jm, k) + F(ekm, i - 1, j, k));
          double epmo = 0.25 * (F(ekm, i, j, k) + F(ekm, i, jm, k) + F(ekm, i + 1, jm, k) + F(ekm, i + 1, j, k));
          T(putout, i, j, k) = T(putout, i, j, k)
            + (epmo * ((F(v0, i + 1, j, k) - F(v0, i, j, k)) * dxi + (F(u0, i + 1, j, k) - F(u0, i + 1, jm, k)) * dyi)
             - emmo * ((F(v0, i, j, k) - F(v0, i - 1, j, k)) * dxi + (F(u0, i, j, k) - F(u0, i, jm, k)) * dyi)) * dxi
            + (F(ekm, i, j, k) * (F(v0, i, jp, k) - F(v0, i, j, k))
             - F(ekm, i, jm, k) * (F(v0, i, j, k) - F(v0, i, jm, k))) * 2. * dy2i
            + (eomp * ((F(v0, i, j, kp) - F(v0, i, j, k)) * M(dzhi, kp) + (F(w0, i, j, kp) - F(w0, i, jm, kp)) * dyi)
             - eomm * ((F(v0, i, j, k) - F(v0, i, j, km)) * M(dzhi, k) + (F(w0, i, j, k) - F(w0, i, jm, k)) * dyi)) * M(dzfi, k);
        } else {
          T(putout, i, j, k) = T(putout, i, j, k)
            + (numol * ((F(v0, i + 1, j, k) - F(v0, i, j, k)) * dxi + (F(u0, i + 1, j, k) - F(u0, i + 1, jm, k)) * dyi)
             - numol * ((F(v0, i, j, k) - F(v0, i - 1, j, k)) * dxi + (F(u0, i, j, k) - F(u0, i, jm, k)) * dyi)) * dxi
            + (numol * (F(v0, i, jp, k) - F(v0, i, j, k))
             - numol * (F(v0, i, j, k) - F(v0, i, jm, k))) * 2. * dy2i
            + (numol * ((F(v0, i, j, kp) - F(v0, i, j, k)) * M(dzhi, kp) + (F(w0, i, j, kp) - F(w0, i, jm, kp)) * dyi)
             - numol * ((F(v0, i, j, k) - F(v0, i, j, km)) * M(dzhi, k) + (F(w0, i, j, k) - F(w0, i, jm, k)) * dyi)) * M(dzfi, k);
        }
      }
    }
  }
}

/* diffw: src/modsubgrid.f90:890-997 */
static void diffw(orc_t *o, double *putout) {
  const double *u0 = o->u0, *v0 = o->v0, *w0 = o->w0, *ekm = o->ekm;
  const double dxi = o->dxi, dyi = o->dyi, numol = o->c.numol;
  const int lles = o->c.lles;
#pragma omp parallel for schedule(static)
  for (int k = 2; k <= o->ktot; k++) {
    const int kp = k + 1, km = k - 1;
    for (int j = 1; j <= o->jtot; j++) {
      const int jp = j + 1, jm = j - 1;
      for (int i = 1; i <= o->itot; i++) {
        if (lles) {
          double emom = (M(dzf, km) * (F(ekm, i, j, k) + F(ekm, i - 1, j, k)) +
                         M(dzf, k) * (F(ekm, i, j, km) + F(ekm, i - 1, j, km))) * M(dzhiq, k);
          double eomm = (M(dzf, km) * (F(ekm, i, j, k) + F(ekm, i, jm, k)) +
                         M(dzf, k) * (F(ekm, i, j, km) + F(ekm, i, jm, km))) * M(dzhiq, k);
          double eopm = (M(dzf, km) * (F(ekm, i, j, k) + F(ekm, i, jp, k)) +
                         M(dzf, k) * (F(ekm, i, j, km) + F(ekm, i, jp, km))) * M(dzhiq, k);
          double epom = (M(dzf, km) * (F(ekm, i, j, k) + F(ekm, i + 1, j, k)) +
                         M(dzf, k) * (F(ekm, i, j, km) + F(ekm, i + 1, j, km))) * M(dzhiq, k);
          T(putout, i, j, k) = T(putout, i, j, k)
            + (epom * ((F(w0, i + 1, j, k) - F(w0, i, j, k)) * dxi + (F(u0, i + 1, j, k) - F(u0, i + 1, j, km)) * M(dzhi, k))
             - emom * ((F(w0, i, j, k) - F(w0, i - 1, j, k)) * dxi + (F(u0, i, j, k) - F(u0, i, j, km)) * M(dzhi, k))) * dxi
            + (eopm * ((F(w0, i, jp, k) - F(w0, i, j, k)) * dyi + (F(v0, i, jp, k) - F(v0, i, jp, km)) * M(dzhi, k))
             - eomm * ((F(w0, i, j, k) - F(w0, i, jm, k)) * dyi + (F(v0, i, j, k) - F(v0, i, j, km)) * M(dzhi, k))) * dyi
            + (F(ekm, i, j, k) * (F(w0, i, j, kp) - F(w0, i, j, k)) * M(dzfi, k)
             - F(ekm, i, j, km) * (F(w0, i, j, k) - F(w0, i, j, km)) * M(dzfi, km)) * 2. * M(dzhi, k);
        } else {
          T(putout, i, j, k) = T(putout, i, j, k)
            + (numol * ((F(w0, i + 1, j, k) - F(w0, i, j, k)) * dxi + (F(u0, i + 1, j, k) - F(u0, i + 1, j, km)) * M(dzhi, k))
             - numol * ((F(w0, i, j, k) - F(w0, i - 1, j, k)) * dxi + (F(u0, i, j, k) - F(u0, i, j, km)) * M(dzhi, k))) * dxi
            + (numol * ((F(w0, i, jp, k) - F(w0, i, j, k)) * dyi + (F(v0, i, jp, k) - F(v0, i, jp, km)) * M(dzhi, k))
             - numol * ((F(w0, i, j, k) - F(w0, i, jm, k)) * dyi + (F(v0, i, j, k) - F(v0, i, j, km)) * M(dzhi, k))) * dyi
            + (numol * (F(w0, i, j, kp) - F(w0, i, j, k)) * M(dzfi, k)
             - numol * (F(w0, i, j, k) - F(w0, i, j, km)) * M(dzfi, km)) * 2. * M(dzhi, k);
        }
      }
    }
  }
}

/* diffc: src/modsubgrid.f90:540-623 on scalar-halo arrays; ekh has the momentum halo */
static void diffc(orc_t *o, const double *putin, double *putout) {
  const double *ekh = o->ekh;
  const double dx2i = o->dx2i, dy2i = o->dy2i;
  const double cekh = o->c.numol * o->c.prandtlmoli;
  const int lles = o->c.lles;
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= o->ktot; k++) {
    const int kp = k + 1, km = k - 1;
    for (int j = 1; j <= o->jtot; j++) {
      const int jp = j + 1, jm = j - 1;
      for (int i = 1; i <= o->itot; i++) {
        const int ip = i + 1, im = i - 1;
        if (lles) {
          ST(putout, i, j, k) = ST(putout, i, j, k)
            + 0.5 * (
                ((F(ekh, ip, j, k) + F(ekh, i, j, k)) * (S(putin, ip, j, k) - S(putin, i, j, k))
               - (F(ekh, i, j, k) + F(ekh, im, j, k)) * (S(putin, i, j, k) - S(putin, im, j, k))) * dx2i
              + ((F(ekh, i, jp, k) + F(ekh, i, j, k)) * (S(putin, i, jp, k) - S(putin, i, j, k))
               - (F(ekh, i, j, k) + F(ekh, i, jm, k)) * (S(putin, i, j, k) - S(putin, i, jm, k))) * dy2i
              + ((M(dzf, kp) * F(ekh, i, j, k) + M(dzf, k) * F(ekh, i, j, kp)) * (S(putin, i, j, kp) - S(putin, i, j, k)) * M(dzh2i, kp)
               - (M(dzf, km) * F(ekh, i, j, k) + M(dzf, k) * F(ekh, i, j, km)) * (S(putin, i, j, k) - S(putin, i, j, km)) * M(dzh2i, k)) * M(dzfi, k));
        } else {
          ST(putout, i, j, k) = ST(putout, i, j, k)
            + ((cekh * (S(putin, ip, j, k) - S(putin, i, j, k)) - cekh * (S(putin, i, j, k) - S(putin, im, j, k))) * dx2i
             + (cekh * (S(putin, i, jp, k) - S(putin, i, j, k)) - cekh * (S(putin, i, j, k) - S(putin, i, jm, k))) * dy2i
             + (cekh * (S(putin, i, j, kp) - S(putin, i, j, k)) * M(dzhi, kp)
              - cekh * (S(putin, i, j, k) - S(putin, i, j, km)) * M(dzhi, k)) * M(dzfi, k));
        }
      }
    }
  }
}

/* subgrid: src/modsubgrid.f90:128-152 */
void orc_subgrid(orc_t *o) {
  orc_closure(o);
  diffu(o, o->up);
  diffv(o, o->vp);
  diffw(o, o->wp);
  if (o->ltempeq) {      /* diffc(ih, jh, kh, thl0, thlp), src/modsubgrid.f90:146 */
    const int hc = o->ihc;
    o->ihc = o->jhc = o->khc = 1;
    diffc(o, o->thl0, o->thlp);
    o->ihc = o->jhc = o->khc = hc;
  }
  for (int n = 0; n < o->nsv; n++) diffc(o, o->sv0 + n * nS(o), o->svp + n * nST(o));
}

/* ------------------------------------------------------------------------- */
/* fillps + bcpup: src/modpois.f90:911-973, src/modboundary.f90:1191-1255,1307-1315 */
void orc_fillps(orc_t *o, double dt, int rk3step) {
  const int I = o->itot, J = o->jtot, K = o->ktot;
  double rk3coef = (rk3step == 0) ? 1. : dt / (4. - (double)rk3step);
  double rk3coefi = 1. / rk3coef;
  double *pup = o->pup, *pvp = o->pvp, *pwp = o->pwp, *p = o->p;
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= K; k++)
    for (int j = 1; j <= J; j++)
      for (int i = 1; i <= I; i++) {
        T(pup, i, j, k) = T(o->up, i, j, k) + F(o->um, i, j, k) * rk3coefi;
        T(pvp, i, j, k) = T(o->vp, i, j, k) + F(o->vm, i, j, k) * rk3coefi;
        T(pwp, i, j, k) = T(o->wp, i, j, k) + F(o->wm, i, j, k) * rk3coefi;
      }
  /* bcpup, freeslip / noslip top */
  for (int j = 1; j <= J; j++)
    for (int i = 1; i <= I; i++) { T(pwp, i, j, 1) = 0.; T(pwp, i, j, K + 1) = 0.; }
  for (int k = 1; k <= K; k++)
    for (int j = 1; j <= J; j++) T(pup, I + 1, j, k) = T(pup, 1, j, k);
  for (int k = 1; k <= K; k++)
    for (int i = 1; i <= I; i++) T(pvp, i, J + 1, k) = T(pvp, i, 1, k);
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= K; k++)
    for (int j = 1; j <= J; j++)
      for (int i = 1; i <= I; i++)
        F(p, i, j, k) = (T(pup, i + 1, j, k) - T(pup, i, j, k)) * o->dxi
                      + (T(pvp, i, j + 1, k) - T(pvp, i, j, k)) * o->dyi
                      + (T(pwp, i, j, k + 1) - T(pwp, i, j, k)) * M(dzfi, k);
}

/* ---- real FFT with FFTW r2c / c2r conventions ---------------------------- */
/* complex radix-2 / generic mixed FFT, sign = -1 forward (FFTW_FORWARD), +1 backward, unnormalised */
static void cfft_rec(int n, int s, const double *in, double *out, int sign, double *scratch) {
  /* in: stride s (complex interleaved), out: contiguous n complex */
  if (n == 1) { out[0] = in[0]; out[1] = in[1]; return; }
  int r = 0;
  if (n % 2 == 0) r = 2; else { for (int q = 3; q * q <= n; q += 2) if (n % q == 0) { r = q; break; } if (!r) r = n; }
  const int m = n / r;
  if (m == 1) {                                      /* prime length: direct DFT */
    for (int k = 0; k < n; k++) {
      double sr = 0, si = 0;
      for (int t = 0; t < n; t++) {
        double ang = sign * 2.0 * M_PI * (double)(((long)k * t) % n) / n;
        double c = cos(ang), sn = sin(ang);
        sr += in[2 * s * t] * c - in[2 * s * t + 1] * sn;
        si += in[2 * s * t] * sn + in[2 * s * t + 1] * c;
      }
      out[2 * k] = sr; out[2 * k + 1] = si;
    }
    return;
  }
  /* decimation in time: r sub-transforms of length m */
  for (int q = 0; q < r; q++) cfft_rec(m, s * r, in + 2 * s * q, scratch + 2 * m * q, sign, out + 2 * m * q);
  for (int k = 0; k < m; k++)
    for (int q2 = 0; q2 < r; q2++) {
      double sr = 0, si = 0;
      const int kk = k + q2 * m;
      for (int q = 0; q < r; q++) {
        double ang = sign * 2.0 * M_PI * (double)(((long)q * kk) % n) / n;
        double c = cos(ang), sn = sin(ang);
        double xr = scratch[2 * (m * q + k)], xi = scratch[2 * (m * q + k) + 1];
        sr += xr * c - xi * sn;
        si += xr * sn + xi * c;
      }
      out[2 * kk] = sr; out[2 * kk + 1] = si;
    }
}

/* iterative power-of-two complex FFT with a twiddle table (fast path for the CPU baseline) */
static void cfft_pow2(int n, double *x, int sign, const double *tw /* (n tws)/2 complex, forward, table of length n * tws */, int tws) {
  for (int i = 1, j = 0; i < n; i++) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { double tr = x[2 * i], ti = x[2 * i + 1]; x[2 * i] = x[2 * j]; x[2 * i + 1] = x[2 * j + 1]; x[2 * j] = tr; x[2 * j + 1] = ti; }
  }
  for (int len = 2; len <= n; len <<= 1) {
    const int half = len >> 1, step = n / len;
    for (int i = 0; i < n; i += len)
      for (int k = 0; k < half; k++) {
        double wr = tw[2 * k * step * tws], wi = sign < 0 ? tw[2 * k * step * tws + 1] : -tw[2 * k * step * tws + 1];
        double *a = x + 2 * (i + k), *b = x + 2 * (i + k + half);
        double tr = b[0] * wr - b[1] * wi, ti = b[0] * wi + b[1] * wr;
        b[0] = a[0] - tr; b[1] = a[1] - ti; a[0] += tr; a[1] += ti;
      }
  }
}

static double *make_tw(int n) { /* exp(-2 pi i k / n), k < n/2 */
  double *tw = (double *)malloc(sizeof(double) * (n > 1 ? n : 2));
  for (int k = 0; k < n / 2; k++) { tw[2 * k] = cos(2.0 * M_PI * k / n); tw[2 * k + 1] = -sin(2.0 * M_PI * k / n); }
  return tw;
}
static int is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

/* forward: real line[n] -> packed [Re0, Re1, Im1, ..., Re(n/2)] * n^-1/2   (modpois.f90:478-490)
 * inverse: packed -> real, c2r unnormalised, * n^-1/2                       (modpois.f90:669-679)
 * work: 4n doubles.  tw: table for n (only used when n is a power of two). */
static void rfft_packed(int n, double *line, int inverse, double *work, const double *tw) {
  const double fac = 1. / sqrt(n * 1.);
  double *z = work, *scr = work + 2 * n;
  if (is_pow2(n) && n >= 4) {
    /* what an r2c / c2r plan does (the reference calls FFTW's, src/modpois.f90:110-111, 481, 676): a real line of n points
     * is a complex FFT of h = n/2 points z[m] = x[2m] + i x[2m+1] plus a split (forward) or merge (inverse) pass.
     * w^k = exp(-2 pi i k / n) = tw[k]. */
    const int h = n / 2;
    if (!inverse) {
      for (int m = 0; m < 2 * h; m++) z[m] = line[m];                 /* (Re, Im) pairs are the line itself */
      cfft_pow2(h, z, -1, tw, 2);
      line[0] = (z[0] + z[1]) * fac;
      line[n - 1] = (z[0] - z[1]) * fac;
      for (int k = 1; k <= h / 2; k++) {
        const double zr = z[2 * k], zi = z[2 * k + 1], cr = z[2 * (h - k)], ci = -z[2 * (h - k) + 1];   /* Zk, conj(Z(h-k)) */
        const double er = 0.5 * (zr + cr), ei = 0.5 * (zi + ci), dr = zr - cr, di = zi - ci;
        const double wr = tw[2 * k], wi = tw[2 * k + 1];
        /* T = (-i/2) w^k D = (0.5 wi, -0.5 wr) * (dr, di) */
        const double ar = 0.5 * wi, ai = -0.5 * wr;
        const double tr = ar * dr - ai * di, ti = ar * di + ai * dr;
        line[2 * k - 1] = (er + tr) * fac; line[2 * k] = (ei + ti) * fac;                               /* X[k] = E + T */
        line[2 * (h - k) - 1] = (er - tr) * fac; line[2 * (h - k)] = -(ei - ti) * fac;                  /* X[h-k] = conj(E - T) */
      }
    } else {
      z[0] = line[0] + line[n - 1]; z[1] = line[0] - line[n - 1];
      for (int k = 1; k <= h / 2; k++) {
        const double xr = line[2 * k - 1], xi = line[2 * k], yr = line[2 * (h - k) - 1], yi = -line[2 * (h - k)];   /* Xk, conj(X(h-k)) */
        const double ar = xr + yr, ai = xi + yi, br = xr - yr, bi = xi - yi;
        const double wr = tw[2 * k], wi = tw[2 * k + 1];
        /* T = i conj(w^k) B = (wi, wr) * (br, bi) */
        const double tr = wi * br - wr * bi, ti = wi * bi + wr * br;
        z[2 * k] = ar + tr; z[2 * k + 1] = ai + ti;                                                     /* Z[k] = A + T */
        z[2 * (h - k)] = ar - tr; z[2 * (h - k) + 1] = -(ai - ti);                                      /* Z[h-k] = conj(A - T) */
      }
      cfft_pow2(h, z, +1, tw, 2);
      for (int m = 0; m < 2 * h; m++) line[m] = z[m] * fac;
    }
    return;
  }
  if (!inverse) {
    for (int i = 0; i < n; i++) { z[2 * i] = line[i]; z[2 * i + 1] = 0.; }
    if (is_pow2(n)) cfft_pow2(n, z, -1, tw, 1);
    else { double *out = (double *)malloc(sizeof(double) * 2 * n); cfft_rec(n, 1, z, out, -1, scr); memcpy(z, out, sizeof(double) * 2 * n); free(out); }
    line[0] = z[0];
    for (int i = 1; i <= n / 2 - 1; i++) { line[2 * i - 1] = z[2 * i]; line[2 * i] = z[2 * i + 1]; }
    line[n - 1] = z[2 * (n / 2)];
    for (int i = 0; i < n; i++) line[i] = line[i] * fac;
  } else {
    z[0] = line[0]; z[1] = 0.;
    for (int i = 1; i <= n / 2 - 1; i++) {
      z[2 * i] = line[2 * i - 1]; z[2 * i + 1] = line[2 * i];
      z[2 * (n - i)] = line[2 * i - 1]; z[2 * (n - i) + 1] = -line[2 * i];   /* Hermitian extension = what c2r implies */
    }
    z[2 * (n / 2)] = line[n - 1]; z[2 * (n / 2) + 1] = 0.;
    if (is_pow2(n)) cfft_pow2(n, z, +1, tw, 1);
    else { double *out = (double *)malloc(sizeof(double) * 2 * n); cfft_rec(n, 1, z, out, +1, scr); memcpy(z, out, sizeof(double) * 2 * n); free(out); }
    for (int i = 0; i < n; i++) line[i] = z[2 * i] * fac;
  }
}

void orc_rfft_packed(int n, double *line, int inverse) {
  double *work = (double *)malloc(sizeof(double) * 4 * n);
  double *tw = make_tw(n);
  rfft_packed(n, line, inverse, work, tw);
  free(work); free(tw);
}

/* solmpj: src/modpois.f90:1107-1166, bxyzrt built on the fly per src/modpois.f90:196-220 */
static void solmpj(orc_t *o, double *x) {
  const int I = o->itot, J = o->jtot, K = o->ktot;
  double *d = o->d;
  const double *a = o->a, *b = o->b, *c = o->cc;
#pragma omp parallel for schedule(static)
  for (int j = 1; j <= J; j++) {
    for (int i = 1; i <= I; i++) {
      double xyzrt = 1.0 * (o->xrt[i] + o->yrt[j] + 0.);           /* rhobf*(xrt+yrt+zrt) */
      double bxyzrt = (xyzrt == 0. && 1 == K) ? o->b_top_D : b[1] + xyzrt;
      double z = 1. / bxyzrt;
      R(d, i, j, 1) = c[1] * z;
      R(x, i, j, 1) = R(x, i, j, 1) * z;
    }
    for (int k = 2; k <= K - 1; k++)
      for (int i = 1; i <= I; i++) {
        double xyzrt = 1.0 * (o->xrt[i] + o->yrt[j] + 0.);
        double bbk = b[k] + xyzrt;
        double z = 1. / (bbk - a[k] * R(d, i, j, k - 1));
        R(d, i, j, k) = c[k] * z;
        R(x, i, j, k) = (R(x, i, j, k) - a[k] * R(x, i, j, k - 1)) * z;
      }
    const double ak = a[K];
    for (int i = 1; i <= I; i++) {
      double xyzrt = 1.0 * (o->xrt[i] + o->yrt[j] + 0.);
      double bbk = (xyzrt == 0.) ? o->b_top_D : b[K] + xyzrt;
      double z = bbk - ak * R(d, i, j, K - 1);
      R(x, i, j, K) = (R(x, i, j, K) - ak * R(x, i, j, K - 1)) / z;
    }
    for (int k = K - 1; k >= 1; k--)
      for (int i = 1; i <= I; i++) R(x, i, j, k) = R(x, i, j, k) - R(d, i, j, k) * R(x, i, j, k + 1);
  }
}

/* poisson, POISS_FFT2D branch, periodic x/y, BCzp=1: src/modpois.f90:440-712.
 * The 8 pencil transposes are identities on a single pencil. */
void orc_poisson_solve(orc_t *o, double *pz) {
  const int I = o->itot, J = o->jtot, K = o->ktot;
  if (!o->twx) o->twx = make_tw(I);
  if (!o->twy) o->twy = make_tw(J);
#pragma omp parallel
  {
    double *work = (double *)malloc(sizeof(double) * 4 * (I > J ? I : J));
    double *line = (double *)malloc(sizeof(double) * (I > J ? I : J));
#pragma omp for schedule(static)
    for (int k = 1; k <= K; k++) {
      for (int j = 1; j <= J; j++) rfft_packed(I, &R(pz, 1, j, k), 0, work, o->twx);       /* :478-490 */
      for (int i = 1; i <= I; i++) {                                                       /* :522-534 */
        for (int j = 1; j <= J; j++) line[j - 1] = R(pz, i, j, k);
        rfft_packed(J, line, 0, work, o->twy);
        for (int j = 1; j <= J; j++) R(pz, i, j, k) = line[j - 1];
      }
    }
    free(work); free(line);
  }
  solmpj(o, pz);                                                                            /* :553 */
#pragma omp parallel
  {
    double *work = (double *)malloc(sizeof(double) * 4 * (I > J ? I : J));
    double *line = (double *)malloc(sizeof(double) * (I > J ? I : J));
#pragma omp for schedule(static)
    for (int k = 1; k <= K; k++) {
      for (int i = 1; i <= I; i++) {                                                       /* :615-625 */
        for (int j = 1; j <= J; j++) line[j - 1] = R(pz, i, j, k);
        rfft_packed(J, line, 1, work, o->twy);
        for (int j = 1; j <= J; j++) R(pz, i, j, k) = line[j - 1];
      }
      for (int j = 1; j <= J; j++) rfft_packed(I, &R(pz, 1, j, k), 1, work, o->twx);       /* :669-679 */
    }
    free(work); free(line);
  }
}

/* tderive + bcp: src/modpois.f90:1001-1105, src/modboundary.f90:1344-1408 */
void orc_tderive(orc_t *o) {
  const int I = o->itot, J = o->jtot, K = o->ktot;
  double *p = o->p, *up = o->up, *vp = o->vp, *wp = o->wp, *pres0 = o->pres0;
  for (int j = 1; j <= J; j++)
    for (int k = 1; k <= K; k++) { F(p, 0, j, k) = F(p, I, j, k); F(p, I + 1, j, k) = F(p, 1, j, k); }
  for (int i = 1; i <= I; i++)
    for (int k = 1; k <= K; k++) { F(p, i, 0, k) = F(p, i, J, k); F(p, i, J + 1, k) = F(p, i, 1, k); }
#pragma omp parallel for schedule(static)
  for (int j = 1; j <= J; j++)
    for (int i = 1; i <= I; i++) {
      T(up, i, j, 1) = T(up, i, j, 1) - (F(p, i, j, 1) - F(p, i - 1, j, 1)) * o->dxi;
      T(vp, i, j, 1) = T(vp, i, j, 1) - (F(p, i, j, 1) - F(p, i, j - 1, 1)) * o->dyi;
      for (int k = 2; k <= K; k++) {
        T(up, i, j, k) = T(up, i, j, k) - (F(p, i, j, k) - F(p, i - 1, j, k)) * o->dxi;
        T(vp, i, j, k) = T(vp, i, j, k) - (F(p, i, j, k) - F(p, i, j - 1, k)) * o->dyi;
        T(wp, i, j, k) = T(wp, i, j, k) - (F(p, i, j, k) - F(p, i, j, k - 1)) * M(dzhi, k);
      }
    }
#pragma omp parallel for schedule(static)
  for (int k = 0; k <= K + 1; k++)
    for (int j = 0; j <= J + 1; j++)
      for (int i = 0; i <= I + 1; i++) F(pres0, i, j, k) = F(pres0, i, j, k) + F(p, i, j, k);
}

/* poisson: src/modpois.f90:419-903 */
void orc_poisson(orc_t *o, double dt, int rk3step) {
  const int I = o->itot, J = o->jtot, K = o->ktot;
  orc_fillps(o, dt, rk3step);
  for (int k = 1; k <= K; k++)
    for (int j = 1; j <= J; j++)
      for (int i = 1; i <= I; i++) R(o->rhs, i, j, k) = F(o->p, i, j, k);                 /* :433 */
  double *pz = (double *)malloc(sizeof(double) * nR(o));
  memcpy(pz, o->rhs, sizeof(double) * nR(o));                                               /* :445 */
  orc_poisson_solve(o, pz);
  for (int k = 1; k <= K; k++)
    for (int j = 1; j <= J; j++)
      for (int i = 1; i <= I; i++) F(o->p, i, j, k) = R(pz, i, j, k);                      /* :707 */
  free(pz);
  orc_tderive(o);
}

/* ------------------------------------------------------------------------- */
/* tstep_update: src/modtstep.f90:49-154 (spinup = .false. branch) */
void orc_tstep_update(orc_t *o, double *dt, double courant, double diffnr, double dtmax,
                      int ladaptive, int *rk3step, double *courtot_out, double *diffnrtot_out) {
  *rk3step = (*rk3step % 3) + 1;
  if (*rk3step != 1) return;
  if (ladaptive) {
    double courtotl = 0., diffnrtotl = 1e-5;
    for (int k = 1; k <= o->ktot; k++)
      for (int j = 1; j <= o->jtot; j++)
        for (int i = 1; i <= o->itot; i++) {
          courtotl = fmax(courtotl, (fabs(F(o->um, i, j, k)) * o->dxi + fabs(F(o->vm, i, j, k)) * o->dyi
                                     + fabs(F(o->wm, i, j, k)) / M(dzh, k)) * (*dt));
          diffnrtotl = fmax(diffnrtotl, fmax(F(o->ekm, i, j, k) * (M(dzh2i, k) + o->dx2i + o->dy2i) * (*dt),
                                             F(o->ekh, i, j, k) * (M(dzh2i, k) + o->dx2i + o->dy2i) * (*dt)));
        }
    if (courtot_out) *courtot_out = courtotl;
    if (diffnrtot_out) *diffnrtot_out = diffnrtotl;
    *dt = fmin(dtmax, fmin((*dt) * courant / courtotl, (*dt) * diffnr / diffnrtotl));
  } else {
    *dt = dtmax;
  }
}

/* tstep_integrate: src/modtstep.f90:171-340 */
void orc_tstep_integrate(orc_t *o, double dt, int rk3step) {
  const int I = o->itot, J = o->jtot, K = o->ktot;
  const double rk3coef = dt / (4. - (double)rk3step);
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= K; k++)
    for (int j = 1; j <= J; j++)
      for (int i = 1; i <= I; i++) {
        F(o->u0, i, j, k) = F(o->um, i, j, k) + rk3coef * T(o->up, i, j, k);
        F(o->v0, i, j, k) = F(o->vm, i, j, k) + rk3coef * T(o->vp, i, j, k);
        F(o->w0, i, j, k) = F(o->wm, i, j, k) + rk3coef * T(o->wp, i, j, k);
        for (int n = 0; n < o->nsv; n++)
          S(o->sv0 + n * nS(o), i, j, k) = S(o->svm + n * nS(o), i, j, k) + rk3coef * ST(o->svp + n * nST(o), i, j, k);
        if (o->ltempeq) F(o->thl0, i, j, k) = F(o->thlm, i, j, k) + rk3coef * T(o->thlp, i, j, k);   /* modtstep.f90:244 */
      }
  if (o->ltempeq) {
    memset(o->thlp, 0, nT(o) * sizeof(double));                         /* :325 */
    if (rk3step == 3) memcpy(o->thlm, o->thl0, nF(o) * sizeof(double)); /* :334 */
  }
  memset(o->up, 0, nT(o) * sizeof(double));
  memset(o->vp, 0, nT(o) * sizeof(double));
  memset(o->wp, 0, nT(o) * sizeof(double));
  if (o->nsv) memset(o->svp, 0, nST(o) * o->nsv * sizeof(double));
  if (rk3step == 3) {
    memcpy(o->um, o->u0, nF(o) * sizeof(double));
    memcpy(o->vm, o->v0, nF(o) * sizeof(double));
    memcpy(o->wm, o->w0, nF(o) * sizeof(double));
    if (o->nsv) memcpy(o->svm, o->sv0, nS(o) * o->nsv * sizeof(double));
  }
}

/* halos: src/modboundary.f90:67-109 -> xm_periodic :508-538, ym_periodic :596-626, xs/ys_periodic */
void orc_halos(orc_t *o) {
  const int I = o->itot, J = o->jtot, K = o->ktot;
  double *mom[8] = {o->u0, o->v0, o->w0, o->um, o->vm, o->wm, o->thl0, o->thlm};   /* xT_periodic / yT_periodic :541-556, :628-647 */
  const int nmom = o->ltempeq ? 8 : 6;
  for (int f = 0; f < nmom; f++) {
    double *a = mom[f];
    for (int m = 1; m <= o->ih; m++)
      for (int k = 1 - o->kh; k <= K + o->kh; k++)
        for (int j = 1 - o->jh; j <= J + o->jh; j++) {
          F(a, 1 - m, j, k) = F(a, I + 1 - m, j, k);
          F(a, I + m, j, k) = F(a, m, j, k);
        }
  }
  for (int n = 0; n < 2 * o->nsv; n++) {
    double *a = (n < o->nsv ? o->sv0 + n * nS(o) : o->svm + (n - o->nsv) * nS(o));
    for (int m = 1; m <= o->ihc; m++)
      for (int k = 1 - o->khc; k <= K + o->khc; k++)
        for (int j = 1 - o->jhc; j <= J + o->jhc; j++) {
          S(a, 1 - m, j, k) = S(a, I + 1 - m, j, k);
          S(a, I + m, j, k) = S(a, m, j, k);
        }
  }
  for (int f = 0; f < nmom; f++) {
    double *a = mom[f];
    for (int m = 1; m <= o->ih; m++)                      /* the reference loops m to ih here too (:603) */
      for (int k = 1 - o->kh; k <= K + o->kh; k++)
        for (int i = 1 - o->ih; i <= I + o->ih; i++) {
          F(a, i, 1 - m, k) = F(a, i, J + 1 - m, k);
          F(a, i, J + m, k) = F(a, i, m, k);
        }
  }
  for (int n = 0; n < 2 * o->nsv; n++) {
    double *a = (n < o->nsv ? o->sv0 + n * nS(o) : o->svm + (n - o->nsv) * nS(o));
    for (int m = 1; m <= o->jhc; m++)
      for (int k = 1 - o->khc; k <= K + o->khc; k++)
        for (int i = 1 - o->ihc; i <= I + o->ihc; i++) {
          S(a, i, 1 - m, k) = S(a, i, J + 1 - m, k);
          S(a, i, J + m, k) = S(a, i, m, k);
        }
  }
}

/* periodic wrap of one momentum-halo array (what exchange_halo_z does on a single periodic pencil) */
static void wrap_mom(orc_t *o, double *a) {
  const int I = o->itot, J = o->jtot, K = o->ktot;
  for (int m = 1; m <= o->ih; m++)
    for (int k = 1 - o->kh; k <= K + o->kh; k++)
      for (int j = 1 - o->jh; j <= J + o->jh; j++) {
        F(a, 1 - m, j, k) = F(a, I + 1 - m, j, k);
        F(a, I + m, j, k) = F(a, m, j, k);
      }
  for (int m = 1; m <= o->jh; m++)
    for (int k = 1 - o->kh; k <= K + o->kh; k++)
      for (int i = 1 - o->ih; i <= I + o->ih; i++) {
        F(a, i, 1 - m, k) = F(a, i, J + 1 - m, k);
        F(a, i, J + m, k) = F(a, i, m, k);
      }
}

/* boundary, periodic x/y subset: src/modboundary.f90:163-204 (+ scalars :238-250 with zero top flux) */
void orc_boundary(orc_t *o) {
  const int I = o->itot, J = o->jtot, K = o->ktot;
  for (int j = 1 - o->jh; j <= J + o->jh; j++)
    for (int i = 1 - o->ih; i <= I + o->ih; i++) { F(o->wm, i, j, 1) = 0.; F(o->w0, i, j, 1) = 0.; }
  if (o->c.BCtopm == 1) {
    fluxtop0(o, o->um); fluxtop0(o, o->u0); fluxtop0(o, o->vm); fluxtop0(o, o->v0);
  } else if (o->c.BCtopm == 2) {
    valuetop(o, o->um, o->c.Uinf); valuetop(o, o->u0, o->c.Uinf);
    valuetop(o, o->vm, o->c.Vinf); valuetop(o, o->v0, o->c.Vinf);
  }
  for (int j = 1 - o->jh; j <= J + o->jh; j++)
    for (int i = 1 - o->ih; i <= I + o->ih; i++) { F(o->w0, i, j, K + 1) = 0.; F(o->wm, i, j, K + 1) = 0.; }
  if (o->ltempeq) {      /* modboundary.f90:208-221 */
    if (o->BCtopT == 1) { fluxtop(o, o->thlm, o->ekh, o->wttop); fluxtop(o, o->thl0, o->ekh, o->wttop); }
    else { valuetop(o, o->thlm, o->thl_top); valuetop(o, o->thl0, o->thl_top); }
  }
  /* fluxtopscal with wsvtop = 0 (modboundary.f90:1521-1537): ghost levels ke+1..ke+khc over the
   * momentum-halo footprint (ib-ih:ie+ih, jb-jh:je+jh) copy level ke */
  for (int n = 0; n < o->nsv; n++) {
    double *s0 = o->sv0 + n * nS(o), *sm = o->svm + n * nS(o);
    for (int m = 1; m <= o->khc; m++)
      for (int j = 1 - o->jh; j <= J + o->jh; j++)
        for (int i = 1 - o->ih; i <= I + o->ih; i++) {
          S(s0, i, j, K + m) = S(s0, i, j, K);
          S(sm, i, j, K + m) = S(sm, i, j, K);
        }
  }
}

/* chkdiv: src/modchecksim.f90:161-203 (+ RMS, our parity metric) */
void orc_chkdiv(orc_t *o, double *divmax, double *divtot, double *divrms) {
  double dmax = 0., dtot = 0., d2 = 0.;
  for (int k = 1; k <= o->ktot; k++)
    for (int j = 1; j <= o->jtot; j++)
      for (int i = 1; i <= o->itot; i++) {
        double div = (F(o->u0, i + 1, j, k) - F(o->u0, i, j, k)) * o->dxi
                   + (F(o->v0, i, j + 1, k) - F(o->v0, i, j, k)) * o->dyi
                   + (F(o->w0, i, j, k + 1) - F(o->w0, i, j, k)) * M(dzfi, k);
        dmax = fmax(dmax, fabs(div));
        dtot = dtot + div * o->dx * o->dy * M(dzf, k);
        d2 += div * div;
      }
  if (divmax) *divmax = dmax;
  if (divtot) *divtot = dtot;
  if (divrms) *divrms = sqrt(d2 / ((double)o->itot * o->jtot * o->ktot));
}

/* randomize_field: src/modstartup.f90:2367-2396 applied to every level k = 1..ktot */
void orc_randomize(orc_t *o, const char *name, int n4, double ampl, int ir) {
  int dims[4];
  double *f = orc_field(o, name, dims);
  if (!f) return;
  const int scal = !strncmp(name, "sv", 2);
  if (scal) f += (size_t)n4 * nS(o);
  const long imm = 134456, ia = 8121, ic = 28411;
  for (int k = 1; k <= o->ktot; k++)
    for (int j = 1; j <= o->jtot; j++)
      for (int i = 1; i <= o->itot; i++) {
        long linear_id = (long)i + (long)o->itot * (long)(j - 1) + (long)o->itot * (long)o->jtot * (long)(k - 1);
        long state = ((long)ir + linear_id) % imm;
        state = (state * ia + ic) % imm;
        double ran = (double)state / (double)imm;
        if (scal) S(f, i, j, k) = S(f, i, j, k) + (ran - 0.5) * 2.0 * ampl;
        else F(f, i, j, k) = F(f, i, j, k) + (ran - 0.5) * 2.0 * ampl;
      }
}

/* one RK3 substep restricted to the in-scope calls of src/program.f90:132-207 */
/* ------------------------------------------------------------------------- */
/* Immersed boundary masking (next tier): solid (src/modibm.f90:748-826), ibmnorm (:697-745,
 * momentum + kappa scalars), diffu/v/w/c_corr (:990-1164) and the mask construction of initibm (:153-192). */
void orc_ibm_set_points(orc_t *o, int kind, int n, const int *ijk) {
  free(o->ibm_pts[kind]);
  o->ibm_pts[kind] = (int *)malloc(sizeof(int) * 3 * (size_t)(n > 0 ? n : 1));
  memcpy(o->ibm_pts[kind], ijk, sizeof(int) * 3 * (size_t)n);
  o->ibm_n[kind] = n;
}
static void wrap_mom(orc_t *o, double *f);
/* initibm :153-192: masks = 1, level kb-kh = 0 (and mask_w(kb) = 0), solid points = 0, then the halo exchange */
void orc_ibm_build_masks(orc_t *o) {
  for (int m = 0; m < 4; m++) {
    if (!o->mask[m]) o->mask[m] = zalloc(nF(o));
    double *mk = o->mask[m];
    for (size_t q = 0; q < nF(o); q++) mk[q] = 1.;
    for (int j = 1 - o->jh; j <= o->jtot + o->jh; j++)
      for (int i = 1 - o->ih; i <= o->itot + o->ih; i++) {
        F(mk, i, j, 0) = 0.;
        if (m == 2) F(mk, i, j, 1) = 0.;
      }
    for (int n = 0; n < o->ibm_n[m]; n++) {
      const int *q = o->ibm_pts[m] + 3 * n;
      F(mk, q[0], q[1], q[2]) = 0.;
    }
    wrap_mom(o, mk);
  }
  o->libm = 1;
}
double *orc_ibm_mask(orc_t *o, int m) { return o->mask[m]; }

/* solid() without mask: var = val, rhs = 0 at the solid points (:762-770) */
static void solid_mom(orc_t *o, int kind, double *var, double *rhs) {
  for (int n = 0; n < o->ibm_n[kind]; n++) {
    const int *q = o->ibm_pts[kind] + 3 * n;
    F(var, q[0], q[1], q[2]) = 0.;
    T(rhs, q[0], q[1], q[2]) = 0.;
  }
}
/* solid() with mask on a scalar-halo array (:772-822): zero-flux attempt = average of the fluid neighbours */
static void solid_scalar(orc_t *o, double *var, double *rhs, double val) {
  const double eps1 = 1.e-10;
  const double *mk = o->mask[3];
  static const int nb[6][3] = {{0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}, {1, 0, 0}, {-1, 0, 0}};   /* order of :781-815 */
  for (int n = 0; n < o->ibm_n[3]; n++) {
    const int *q = o->ibm_pts[3] + 3 * n;
    const int i = q[0], j = q[1], k = q[2];
    S(var, i, j, k) = val;
    ST(rhs, i, j, k) = 0.;
    double count = 0.;
    for (int d = 0; d < 6; d++) {
      const int ii = i + nb[d][0], jj = j + nb[d][1], kk = k + nb[d][2];
      if (fabs(F(mk, ii, jj, kk) - 1.) < eps1) {
        count = count + 1.;
        S(var, i, j, k) = S(var, i, j, k) + S(var, ii, jj, kk);
        ST(rhs, i, j, k) = ST(rhs, i, j, k) + ST(rhs, ii, jj, kk);
      }
    }
    if (count > 0.) {
      S(var, i, j, k) = (S(var, i, j, k) - val) / count;
      ST(rhs, i, j, k) = ST(rhs, i, j, k) / count;
    }
  }
}
/* advecc2nd_corr_liberal :936-987 on a momentum-halo scalar (thl with iadv_thl = cd2) */
static void advecc2nd_corr_liberal(orc_t *o, const double *var, double *rhs) {
  const double eps1 = 1.e-10;
  const double *mk = o->mask[3], *u0 = o->u0, *v0 = o->v0, *w0 = o->w0;
  const double dxi5 = o->dxi5, dyi5 = o->dyi5;
  for (int n = 0; n < o->ibm_n[7]; n++) {
    const int *q = o->ibm_pts[7] + 3 * n;
    const int i = q[0], j = q[1], k = q[2];
    if (fabs(F(mk, i + 1, j, k)) < eps1)
      T(rhs, i, j, k) = T(rhs, i, j, k) + F(u0, i + 1, j, k) * (F(var, i + 1, j, k) + F(var, i, j, k)) * dxi5
                                        - F(u0, i + 1, j, k) * (F(var, i, j, k) + F(var, i, j, k)) * dxi5;
    if (fabs(F(mk, i - 1, j, k)) < eps1)
      T(rhs, i, j, k) = T(rhs, i, j, k) - F(u0, i, j, k) * (F(var, i - 1, j, k) + F(var, i, j, k)) * dxi5
                                        + F(u0, i, j, k) * (F(var, i, j, k) + F(var, i, j, k)) * dxi5;
    if (fabs(F(mk, i, j + 1, k)) < eps1)
      T(rhs, i, j, k) = T(rhs, i, j, k) + F(v0, i, j + 1, k) * (F(var, i, j + 1, k) + F(var, i, j, k)) * dyi5
                                        - F(v0, i, j + 1, k) * (F(var, i, j, k) + F(var, i, j, k)) * dyi5;
    if (fabs(F(mk, i, j - 1, k)) < eps1)
      T(rhs, i, j, k) = T(rhs, i, j, k) - F(v0, i, j, k) * (F(var, i, j - 1, k) + F(var, i, j, k)) * dyi5
                                        + F(v0, i, j, k) * (F(var, i, j, k) + F(var, i, j, k)) * dyi5;
    if (fabs(F(mk, i, j, k + 1)) < eps1)
      T(rhs, i, j, k) = T(rhs, i, j, k) + F(w0, i, j, k + 1) * (F(var, i, j, k + 1) * M(dzf, k) + F(var, i, j, k) * M(dzf, k + 1)) * M(dzhi, k + 1) * M(dzfi5, k)
                                        - F(w0, i, j, k + 1) * (F(var, i, j, k) * M(dzf, k) + F(var, i, j, k) * M(dzf, k + 1)) * M(dzhi, k + 1) * M(dzfi5, k);
    if (fabs(F(mk, i, j, k - 1)) < eps1)
      T(rhs, i, j, k) = T(rhs, i, j, k) - F(w0, i, j, k) * (F(var, i, j, k - 1) * M(dzf, k) + F(var, i, j, k) * M(dzf, k - 1)) * M(dzhi, k) * M(dzfi5, k)
                                        + F(w0, i, j, k) * (F(var, i, j, k) * M(dzf, k) + F(var, i, j, k) * M(dzf, k - 1)) * M(dzhi, k) * M(dzfi5, k);
  }
}
/* ibmnorm :697-745 (momentum + temperature + scalars; kappa scalars need no advecc2nd correction) */
void orc_ibmnorm(orc_t *o) {
  if (!o->libm) return;
  solid_mom(o, 0, o->um, o->up);
  solid_mom(o, 1, o->vm, o->vp);
  solid_mom(o, 2, o->wm, o->wp);
  if (o->ltempeq) {      /* :714-722: solid(.., thlm, thlp, sum(thl0av dzf)/zh(ke+1), .., mask_c) + advecc2nd_corr_liberal */
    const int hc = o->ihc, K = o->ktot;
    double val = 0.;
    for (int k = 1; k <= K; k++) val += o->thl0av[k - 1] * M(dzf, k);
    val = val / M(zh, K + 1);
    o->ihc = o->jhc = o->khc = 1;
    solid_scalar(o, o->thlm, o->thlp, val);
    o->ihc = o->jhc = o->khc = hc;
    advecc2nd_corr_liberal(o, o->thl0, o->thlp);
  }
  for (int n = 0; n < o->nsv; n++) solid_scalar(o, o->svm + n * nS(o), o->svp + n * nST(o), 0.);
}
/* diffu_corr / diffv_corr / diffw_corr / diffc_corr :990-1164: cancel the subgrid flux through solid neighbours */
void orc_ibm_diffcorr(orc_t *o) {
  if (!o->libm) return;
  const double eps1 = 1.e-10;
  const double *ekm = o->ekm, *ekh = o->ekh, *u0 = o->u0, *v0 = o->v0, *w0 = o->w0;
  const double dx2i = o->dx2i, dy2i = o->dy2i;
  for (int n = 0; n < o->ibm_n[4]; n++) {
    const int *q = o->ibm_pts[4] + 3 * n;
    const int i = q[0], j = q[1], k = q[2];
    const double *mk = o->mask[0];
    if (fabs(F(mk, i, j + 1, k)) < eps1) {
      const double empo = 0.25 * ((F(ekm, i, j, k) + F(ekm, i, j + 1, k)) + (F(ekm, i - 1, j, k) + F(ekm, i - 1, j + 1, k)));
      T(o->up, i, j, k) = T(o->up, i, j, k) - empo * (F(u0, i, j + 1, k) - F(u0, i, j, k)) * dy2i;
    }
    if (fabs(F(mk, i, j - 1, k)) < eps1) {
      const double emmo = 0.25 * ((F(ekm, i, j, k) + F(ekm, i, j - 1, k)) + (F(ekm, i - 1, j - 1, k) + F(ekm, i - 1, j, k)));
      T(o->up, i, j, k) = T(o->up, i, j, k) + emmo * (F(u0, i, j, k) - F(u0, i, j - 1, k)) * dy2i;
    }
    if (fabs(F(mk, i, j, k + 1)) < eps1) {
      const double emop = (M(dzf, k + 1) * (F(ekm, i, j, k) + F(ekm, i - 1, j, k)) +
                           M(dzf, k) * (F(ekm, i, j, k + 1) + F(ekm, i - 1, j, k + 1))) * M(dzhiq, k + 1);
      T(o->up, i, j, k) = T(o->up, i, j, k) - emop * (F(u0, i, j, k + 1) - F(u0, i, j, k)) * M(dzhi, k + 1) * M(dzfi, k);
    }
    if (fabs(F(mk, i, j, k - 1)) < eps1) {
      const double emom = (M(dzf, k - 1) * (F(ekm, i, j, k) + F(ekm, i - 1, j, k)) +
                           M(dzf, k) * (F(ekm, i, j, k - 1) + F(ekm, i - 1, j, k - 1))) * M(dzhiq, k);
      T(o->up, i, j, k) = T(o->up, i, j, k) + emom * (F(u0, i, j, k) - F(u0, i, j, k - 1)) * M(dzhi, k) * M(dzfi, k);
    }
  }
  for (int n = 0; n < o->ibm_n[5]; n++) {
    const int *q = o->ibm_pts[5] + 3 * n;
    const int i = q[0], j = q[1], k = q[2];
    const double *mk = o->mask[1];
    if (fabs(F(mk, i + 1, j, k)) < eps1) {
      const double epmo = 0.25 * (F(ekm, i, j, k) + F(ekm, i, j - 1, k) + F(ekm, i + 1, j - 1, k) + F(ekm, i + 1, j, k));
      T(o->vp, i, j, k) = T(o->vp, i, j, k) - epmo * (F(v0, i + 1, j, k) - F(v0, i, j, k)) * dx2i;
    }
    if (fabs(F(mk, i - 1, j, k)) < eps1) {
      const double emmo = 0.25 * (F(ekm, i, j, k) + F(ekm, i, j - 1, k) + F(ekm, i - 1, j - 1, k) + F(ekm, i - 1, j, k));
      T(o->vp, i, j, k) = T(o->vp, i, j, k) + emmo * (F(v0, i, j, k) - F(v0, i - 1, j, k)) * dx2i;
    }
    if (fabs(F(mk, i, j, k + 1)) < eps1) {
      const double eomp = (M(dzf, k + 1) * (F(ekm, i, j, k) + F(ekm, i, j - 1, k)) +
                           M(dzf, k) * (F(ekm, i, j, k + 1) + F(ekm, i, j - 1, k + 1))) * M(dzhiq, k + 1);
      T(o->vp, i, j, k) = T(o->vp, i, j, k) - eomp * (F(v0, i, j, k + 1) - F(v0, i, j, k)) * M(dzhi, k + 1) * M(dzfi, k);
    }
    if (fabs(F(mk, i, j, k - 1)) < eps1) {
      const double eomm = (M(dzf, k - 1) * (F(ekm, i, j, k) + F(ekm, i, j - 1, k)) +
                           M(dzf, k) * (F(ekm, i, j, k - 1) + F(ekm, i, j - 1, k - 1))) * M(dzhiq, k);
      T(o->vp, i, j, k) = T(o->vp, i, j, k) + eomm * (F(v0, i, j, k) - F(v0, i, j, k - 1)) * M(dzhi, k) * M(dzfi, k);
    }
  }
  for (int n = 0; n < o->ibm_n[6]; n++) {
    const int *q = o->ibm_pts[6] + 3 * n;
    const int i = q[0], j = q[1], k = q[2];
    const double *mk = o->mask[2];
    if (fabs(F(mk, i + 1, j, k)) < eps1) {
      const double epom = (M(dzf, k - 1) * (F(ekm, i, j, k) + F(ekm, i + 1, j, k)) +
                           M(dzf, k) * (F(ekm, i, j, k - 1) + F(ekm, i + 1, j, k - 1))) * M(dzhiq, k);
      T(o->wp, i, j, k) = T(o->wp, i, j, k) - epom * (F(w0, i + 1, j, k) - F(w0, i, j, k)) * dx2i;
    }
    if (fabs(F(mk, i - 1, j, k)) < eps1) {
      const double emom = (M(dzf, k - 1) * (F(ekm, i, j, k) + F(ekm, i - 1, j, k)) +
                           M(dzf, k) * (F(ekm, i, j, k - 1) + F(ekm, i - 1, j, k - 1))) * M(dzhiq, k);
      T(o->wp, i, j, k) = T(o->wp, i, j, k) + emom * (F(w0, i, j, k) - F(w0, i - 1, j, k)) * dx2i;
    }
    if (fabs(F(mk, i, j + 1, k)) < eps1) {
      const double eopm = (M(dzf, k - 1) * (F(ekm, i, j, k) + F(ekm, i, j + 1, k)) +
                           M(dzf, k) * (F(ekm, i, j, k - 1) + F(ekm, i, j + 1, k - 1))) * M(dzhiq, k);
      T(o->wp, i, j, k) = T(o->wp, i, j, k) - eopm * (F(w0, i, j + 1, k) - F(w0, i, j, k)) * dy2i;
    }
    if (fabs(F(mk, i, j - 1, k)) < eps1) {
      const double eomm = (M(dzf, k - 1) * (F(ekm, i, j, k) + F(ekm, i, j - 1, k)) +
                           M(dzf, k) * (F(ekm, i, j, k - 1) + F(ekm, i, j - 1, k - 1))) * M(dzhiq, k);
      T(o->wp, i, j, k) = T(o->wp, i, j, k) + eomm * (F(w0, i, j, k) - F(w0, i, j - 1, k)) * dy2i;
    }
  }
  const int hc_save = o->ihc;
  for (int s4 = (o->ltempeq ? -1 : 0); s4 < o->nsv; s4++) {   /* s4 = -1: diffc_corr(thl0, thlp, ih, jh, kh), :1225 */
    const double *var = s4 < 0 ? o->thl0 : o->sv0 + s4 * nS(o);
    double *rhs = s4 < 0 ? o->thlp : o->svp + s4 * nST(o);
    const double *mk = o->mask[3];
    o->ihc = o->jhc = o->khc = s4 < 0 ? 1 : hc_save;
    for (int n = 0; n < o->ibm_n[7]; n++) {
      const int *q = o->ibm_pts[7] + 3 * n;
      const int i = q[0], j = q[1], k = q[2];
      if (fabs(F(mk, i + 1, j, k)) < eps1)
        ST(rhs, i, j, k) = ST(rhs, i, j, k) - 0.5 * (F(ekh, i + 1, j, k) + F(ekh, i, j, k)) * (S(var, i + 1, j, k) - S(var, i, j, k)) * dx2i;
      if (fabs(F(mk, i - 1, j, k)) < eps1)
        ST(rhs, i, j, k) = ST(rhs, i, j, k) + 0.5 * (F(ekh, i, j, k) + F(ekh, i - 1, j, k)) * (S(var, i, j, k) - S(var, i - 1, j, k)) * dx2i;
      if (fabs(F(mk, i, j + 1, k)) < eps1)
        ST(rhs, i, j, k) = ST(rhs, i, j, k) - 0.5 * (F(ekh, i, j + 1, k) + F(ekh, i, j, k)) * (S(var, i, j + 1, k) - S(var, i, j, k)) * dy2i;
      if (fabs(F(mk, i, j - 1, k)) < eps1)
        ST(rhs, i, j, k) = ST(rhs, i, j, k) + 0.5 * (F(ekh, i, j, k) + F(ekh, i, j - 1, k)) * (S(var, i, j, k) - S(var, i, j - 1, k)) * dy2i;
      if (fabs(F(mk, i, j, k + 1)) < eps1)
        ST(rhs, i, j, k) = ST(rhs, i, j, k) - 0.5 * (M(dzf, k + 1) * F(ekh, i, j, k) + M(dzf, k) * F(ekh, i, j, k + 1)) *
                                                  (S(var, i, j, k + 1) - S(var, i, j, k)) * M(dzh2i, k + 1) * M(dzfi, k);
      if (fabs(F(mk, i, j, k - 1)) < eps1)
        ST(rhs, i, j, k) = ST(rhs, i, j, k) + 0.5 * (M(dzf, k - 1) * F(ekh, i, j, k) + M(dzf, k) * F(ekh, i, j, k - 1)) *
                                                  (S(var, i, j, k) - S(var, i, j, k - 1)) * M(dzh2i, k) * M(dzfi, k);
    }
  }
  o->ihc = o->jhc = o->khc = hc_save;
}

/* forces, neutral branch: src/modforces.f90:88-125 */
void orc_set_forcing(orc_t *o, const double *dpdxl, const double *dpdyl) {
  const int n = o->ktot + o->kh;
  if (!o->dpdxl) { o->dpdxl = zalloc(n); o->dpdyl = zalloc(n); }
  memcpy(o->dpdxl, dpdxl, n * sizeof(double));
  memcpy(o->dpdyl, dpdyl, n * sizeof(double));
  o->has_forcing = 1;
}
void orc_forces(orc_t *o) {
  if (!o->has_forcing && !o->ltempeq) return;
  const int I = o->itot, J = o->jtot, K = o->ktot;
  if (!o->dpdxl) { o->dpdxl = zalloc(K + o->kh); o->dpdyl = zalloc(K + o->kh); }
  for (int k = 2; k <= K; k++)
    for (int j = 1; j <= J; j++)
      for (int i = 1; i <= I; i++) {
        T(o->up, i, j, k) = T(o->up, i, j, k) - o->dpdxl[k - 1];
        T(o->vp, i, j, k) = T(o->vp, i, j, k) - o->dpdyl[k - 1];
        if (o->lbuoyancy)    /* src/modforces.f90:78 */
          T(o->wp, i, j, k) = T(o->wp, i, j, k) + o->grav * (T(o->thv0h, i, j, k) - o->thvh[k - 1]) / o->thvh[k - 1];
      }
  if (o->ltempeq)            /* radiative heating, :103-109 */
    for (int k = 1; k <= K; k++)
      for (int j = 1; j <= J; j++)
        for (int i = 1; i <= I; i++) T(o->thlp, i, j, k) = T(o->thlp, i, j, k) + o->thlpcar[k - 1];
  for (int j = 1; j <= J; j++)
    for (int i = 1; i <= I; i++) {
      T(o->up, i, j, 1) = T(o->up, i, j, 1) - o->dpdxl[0];
      T(o->vp, i, j, 1) = T(o->vp, i, j, 1) - o->dpdyl[0];
      T(o->wp, i, j, 1) = 0.0;
    }
}

/* ---- bottom -> wfmneutral(..., 91) ----------------------------------------------------------------------------
 * src/modibm.f90:1998-2100 (lbottom branch: BCbotm = 3 -> wfmneutral; nsv > 0, BCbots = 1: zero-flux scalar
 * correction :2077-2091) and src/modwallfunctions.f90:307-349.  dxf = dx, dxhi = dxi (x is uniform). */
void orc_set_bottom(orc_t *o, int lbottom, int BCbotm, int BCbots, double z0, double fkar) {
  o->lbottom = lbottom; o->BCbotm = BCbotm; o->BCbots = BCbots; o->z0 = z0; o->fkar = fkar;
  if (!o->momfluxb) o->momfluxb = zalloc(nF(o));
}
double *orc_momfluxb(orc_t *o) { return o->momfluxb; }
/* unom / unoh: src/modwallfunctions.f90:226-260 / :176-223 (Uno 1995 stability functions) */
static double unom(double logdz, double logzh, double sqdz, double Ribl0, double fkar2, double prandtlturb) {
  const double b1 = 9.4, b2 = 4.7, dm = 7.4, dh = 5.3;
  double Fm, Fh, cm, ch;
  if (Ribl0 > 0.) { Fm = 1. / ((1. + b2 * Ribl0) * (1. + b2 * Ribl0)); Fh = Fm; }
  else {
    cm = (dm * fkar2) / (logdz * logdz) * b1 * sqdz;
    ch = (dh * fkar2) / (logdz * logdz) * b1 * sqdz;
    Fm = 1. - (b1 * Ribl0) / (1. + cm * sqrt(fabs(Ribl0)));
    Fh = 1. - (b1 * Ribl0) / (1. + ch * sqrt(fabs(Ribl0)));
  }
  const double M = prandtlturb * logdz * sqrt(Fm) / Fh;
  const double Ribl1 = Ribl0 - Ribl0 * prandtlturb * logzh / (prandtlturb * logzh + M);
  if (Ribl1 > 0.) Fm = 1. / ((1. + b2 * Ribl1) * (1. + b2 * Ribl1));
  else {
    cm = (dm * fkar2) / (logdz * logdz) * b1 * sqdz;
    Fm = 1. - (b1 * Ribl1) / (1. + cm * sqrt(fabs(Ribl1)));
  }
  return fkar2 / (logdz * logdz) * Fm;
}
static double unoh(double logdz, double logzh, double sqdz, double utangInt, double dT, double Ribl0, double fkar2, double prandtlturb) {
  const double b1 = 9.4, b2 = 4.7, dm = 7.4, dh = 5.3;
  double Fm, Fh, cm, ch;
  if (Ribl0 > 0.) { Fm = 1. / ((1. + b2 * Ribl0) * (1. + b2 * Ribl0)); Fh = Fm; }
  else {
    cm = (dm * fkar2) / (logdz * logdz) * b1 * sqdz;
    ch = (dh * fkar2) / (logdz * logdz) * b1 * sqdz;
    Fm = 1. - (b1 * Ribl0) / (1. + cm * sqrt(fabs(Ribl0)));
    Fh = 1. - (b1 * Ribl0) / (1. + ch * sqrt(fabs(Ribl0)));
  }
  double M = prandtlturb * logdz * sqrt(Fm) / Fh;
  const double Ribl1 = Ribl0 - Ribl0 * prandtlturb * logzh / (prandtlturb * logzh + M);
  if (Ribl1 > 0.) { Fm = 1. / ((1. + b2 * Ribl1) * (1. + b2 * Ribl1)); Fh = Fm; }
  else {
    cm = (dm * fkar2) / (logdz * logdz) * b1 * sqdz;
    ch = (dh * fkar2) / (logdz * logdz) * b1 * sqdz;
    Fm = 1. - (b1 * Ribl1) / (1. + cm * sqrt(fabs(Ribl1)));
    Fh = 1. - (b1 * Ribl1) / (1. + ch * sqrt(fabs(Ribl1)));
  }
  M = prandtlturb * logdz * sqrt(Fm) / Fh;
  const double dTrough = dT * 1. / (prandtlturb * logzh / M + 1.);
  const double octh = sqrt(utangInt) * fkar2 / (logdz * logdz) * Fh / prandtlturb;
  return octh * dTrough;
}
void orc_set_wfuno(orc_t *o, double z0h, double prandtlturb, double grav, double thls, double tcell) {
  o->wf_z0h = z0h; o->wf_prandtlturb = prandtlturb; o->wf_grav = grav; o->wf_twall = thls; o->wf_tcell = tcell;
}
void orc_bottom(orc_t *o) {
  if (!o->lbottom) return;
  const int I = o->itot, J = o->jtot;
  const double *u0 = o->u0, *v0 = o->v0, *ekm = o->ekm, *ekh = o->ekh;
  if (o->BCbotm == 2) {   /* wfuno(.., 91): src/modwallfunctions.f90:79-128; Tcell = thl0 (uniform without temperature equation) */
    const int k = 1, km = 0;
    const double fkar2 = o->fkar * o->fkar, umin = 0.0001, Twall = o->wf_twall, grav = o->wf_grav, pt = o->wf_prandtlturb;
    const double delta = 0.5 * M(dzf, k);
    const double logdz = log(delta / o->z0), logzh = log(o->z0 / o->wf_z0h), sqdz = sqrt(delta / o->z0);
    const double dx = o->dx, dxhi = o->dxi;
#define TC(i, j) (o->ltempeq ? F(o->thl0, i, j, k) : o->wf_tcell)
    for (int j = 1; j <= J; j++)
      for (int i = 1; i <= I; i++) {
        const double utang1Int = F(u0, i, j, k);
        const double utang2Int = (F(v0, i, j, k) + F(v0, i - 1, j, k) + F(v0, i, j + 1, k) + F(v0, i - 1, j + 1, k)) * 0.25;
        const double utangInt = fmax(umin, (utang1Int * utang1Int + utang2Int * utang2Int));
        const double dT = ((TC(i, j) + TC(i - 1, j)) - (Twall + Twall)) * 0.5;
        const double Ribl0 = grav * delta * dT * 2 / ((Twall + Twall) * utangInt);
        const double ctm = unom(logdz, logzh, sqdz, Ribl0, fkar2, pt);
        const double dummy = fabs(utang1Int) * sqrt(utangInt) * ctm;
        const double bcmomflux = copysign(dummy, utang1Int);
        F(o->momfluxb, i, j, k) = F(o->momfluxb, i, j, k) + bcmomflux * M(dzfi, k);
        const double emom = (M(dzf, km) * (F(ekm, i, j, k) * dx + F(ekm, i - 1, j, k) * dx) +
                             M(dzf, k) * (F(ekm, i, j, km) * dx + F(ekm, i - 1, j, km) * dx)) * dxhi * M(dzhiq, k);
        T(o->up, i, j, k) = T(o->up, i, j, k) + (F(u0, i, j, k) - F(u0, i, j, km)) * emom * M(dzhi, k) * M(dzfi, k) - bcmomflux * M(dzfi, k);
      }
    for (int j = 1; j <= J; j++)
      for (int i = 1; i <= I; i++) {
        const double utang1Int = (F(u0, i, j, k) + F(u0, i, j - 1, k) + F(u0, i + 1, j - 1, k) + F(u0, i + 1, j, k)) * 0.25;
        const double utang2Int = F(v0, i, j, k);
        const double utangInt = fmax(umin, (utang1Int * utang1Int + utang2Int * utang2Int));
        const double dT = ((TC(i, j) + TC(i, j - 1)) - (Twall + Twall)) * 0.5;
        const double Ribl0 = grav * delta * dT * 2 / ((Twall + Twall) * utangInt);
        const double ctm = unom(logdz, logzh, sqdz, Ribl0, fkar2, pt);
        const double dummy = fabs(utang2Int) * sqrt(utangInt) * ctm;
        const double bcmomflux = copysign(dummy, utang2Int);
        F(o->momfluxb, i, j, k) = F(o->momfluxb, i, j, k) + bcmomflux * M(dzfi, k);
        const double eomm = (M(dzf, km) * (F(ekm, i, j, k) + F(ekm, i, j - 1, k)) + M(dzf, k) * (F(ekm, i, j, km) + F(ekm, i, j - 1, km))) * M(dzhiq, k);
        T(o->vp, i, j, k) = T(o->vp, i, j, k) + (F(v0, i, j, k) - F(v0, i, j, km)) * eomm * M(dzhi, k) * M(dzfi, k) - bcmomflux * M(dzfi, k);
      }
#undef TC
  }
  if (o->BCbotm == 3) {
    const int k = 1, km = 0;
    const double fkar2 = o->fkar * o->fkar, umin = 0.0001;
    const double delta = 0.5 * M(dzf, k);
    const double lg = log(delta / o->z0);
    const double logdz2 = lg * lg;
    const double dx = o->dx, dxhi = o->dxi;
    for (int j = 1; j <= J; j++)       /* u component, modwallfunctions.f90:318-332 */
      for (int i = 1; i <= I; i++) {
        const double utang1Int = F(u0, i, j, k);
        const double utang2Int = (F(v0, i, j, k) + F(v0, i - 1, j, k) + F(v0, i, j + 1, k) + F(v0, i - 1, j + 1, k)) * 0.25;
        const double utangInt = fmax(umin, (utang1Int * utang1Int + utang2Int * utang2Int));
        const double ctm = fkar2 / (logdz2);
        const double dummy = fabs(utang1Int) * sqrt(utangInt) * ctm;
        const double bcmomflux = copysign(dummy, utang1Int);
        F(o->momfluxb, i, j, k) = F(o->momfluxb, i, j, k) + bcmomflux * M(dzfi, k);
        const double emom = (M(dzf, km) * (F(ekm, i, j, k) * dx + F(ekm, i - 1, j, k) * dx) +
                             M(dzf, k) * (F(ekm, i, j, km) * dx + F(ekm, i - 1, j, km) * dx)) * dxhi * M(dzhiq, k);
        T(o->up, i, j, k) = T(o->up, i, j, k) + (F(u0, i, j, k) - F(u0, i, j, km)) * emom * M(dzhi, k) * M(dzfi, k) - bcmomflux * M(dzfi, k);
      }
    for (int j = 1; j <= J; j++)       /* v component, :334-347 */
      for (int i = 1; i <= I; i++) {
        const double utang1Int = (F(u0, i, j, k) + F(u0, i, j - 1, k) + F(u0, i + 1, j - 1, k) + F(u0, i + 1, j, k)) * 0.25;
        const double utang2Int = F(v0, i, j, k);
        const double utangInt = fmax(umin, (utang1Int * utang1Int + utang2Int * utang2Int));
        const double ctm = fkar2 / (logdz2);
        const double dummy = fabs(utang2Int) * sqrt(utangInt) * ctm;
        const double bcmomflux = copysign(dummy, utang2Int);
        F(o->momfluxb, i, j, k) = F(o->momfluxb, i, j, k) + bcmomflux * M(dzfi, k);
        const double eomm = (M(dzf, km) * (F(ekm, i, j, k) + F(ekm, i, j - 1, k)) + M(dzf, k) * (F(ekm, i, j, km) + F(ekm, i, j - 1, km))) * M(dzhiq, k);
        T(o->vp, i, j, k) = T(o->vp, i, j, k) + (F(v0, i, j, k) - F(v0, i, j, km)) * eomm * M(dzhi, k) * M(dzfi, k) - bcmomflux * M(dzfi, k);
      }
  }
  if (o->ltempeq && o->BCbotT == 2) {  /* wfuno(.., 92): fixed wall temperature, src/modwallfunctions.f90:131-161 */
    const int k = 1;
    const double fkar2 = o->fkar * o->fkar, umin = 0.0001, Twall = o->wf_twall, grav = o->wf_grav, pt = o->wf_prandtlturb;
    const double delta = M(dzf, k) * 0.5;
    const double logdz = log(delta / o->z0), logzh = log(o->z0 / o->wf_z0h), sqdz = sqrt(delta / o->z0);
    for (int j = 1; j <= J; j++)
      for (int i = 1; i <= I; i++) {
        const double utang1Int = (F(u0, i, j, k) + F(u0, i + 1, j, k)) * 0.5;
        const double utang2Int = (F(v0, i, j, k) + F(v0, i, j + 1, k)) * 0.5;
        const double utangInt = fmax(umin, (utang1Int * utang1Int + utang2Int * utang2Int));
        const double dT = (F(o->thl0, i, j, k) - Twall);
        const double Ribl0 = grav * delta * dT / (Twall * utangInt);
        const double bcTflux = unoh(logdz, logzh, sqdz, utangInt, dT, Ribl0, fkar2, pt);
        T(o->thlp, i, j, k) = T(o->thlp, i, j, k) + 0.5 * (M(dzf, k - 1) * F(ekh, i, j, k) + M(dzf, k) * F(ekh, i, j, k - 1)) *
                              (F(o->thl0, i, j, k) - F(o->thl0, i, j, k - 1)) * M(dzh2i, k) * M(dzfi, k) - bcTflux * M(dzfi, k);
      }
  }
  if (o->ltempeq && o->BCbotT == 1) {  /* fixed-flux bottom for temperature, modibm.f90:2033-2046 */
    const int kb = 1;
    for (int j = 1; j <= J; j++)
      for (int i = 1; i <= I; i++)
        T(o->thlp, i, j, kb) = T(o->thlp, i, j, kb) + (0.5 * (M(dzf, kb - 1) * F(ekh, i, j, kb) + M(dzf, kb) * F(ekh, i, j, kb - 1)) *
                                                      (F(o->thl0, i, j, kb) - F(o->thl0, i, j, kb - 1)) * M(dzh2i, kb) - o->wtsurf) * M(dzfi, kb);
  }
  if (o->nsv > 0 && o->BCbots == 1) {  /* modibm.f90:2077-2091 */
    const int kb = 1;
    for (int n = 0; n < o->nsv; n++) {
      const double *sv0 = o->sv0 + (size_t)n * nS(o);
      double *svp = o->svp + (size_t)n * nST(o);
      for (int j = 1; j <= J; j++)
        for (int i = 1; i <= I; i++)
          ST(svp, i, j, kb) = ST(svp, i, j, kb) + (0.5 * (M(dzf, kb - 1) * F(ekh, i, j, kb) + M(dzf, kb) * F(ekh, i, j, kb - 1)) *
                                                   (S(sv0, i, j, kb) - S(sv0, i, j, kb - 1)) * M(dzh2i, kb) + 0.) * M(dzfi, kb);
    }
  }
}

/* ---- temperature: set-up and thermodynamics (dry) ---------------------------------------------------------------
 * src/modthermodynamics.f90:55-121 with lmoist = .false.: diagfld's slab mean of thl0 (:270, avexy_ibm with IIc),
 * calc_halflev (:508-526), calthv (:213-235), thvh = avexy_ibm(thv0h, IIw) with the kb / kb+1 overrides (:76-90).
 * Not restated: the hydrostatic pressure / exner / density profiles (fromztop, :338-420) and thvf; nothing on the
 * dry path reads them (they feed the moist thermodynamics and the statistics). */
void orc_set_thermo(orc_t *o, int lbuoyancy, double grav, double thls, int BCtopT, double wttop, double thl_top,
                    int BCbotT, double wtsurf, const double *thlpcar) {
  const int K = o->ktot;
  if (!o->thl0) {
    o->thl0 = zalloc(nF(o)); o->thlm = zalloc(nF(o)); o->thl0h = zalloc(nF(o));
    o->thlp = zalloc(nT(o)); o->thv0h = zalloc(nT(o)); o->dthvdz = zalloc(nT(o));
    o->thl0av = zalloc(K + 1); o->thvh = zalloc(K + 1); o->thlpcar = zalloc(K + 1);
  }
  o->ltempeq = 1; o->lbuoyancy = lbuoyancy; o->grav = grav; o->thls = thls;
  o->BCtopT = BCtopT; o->wttop = wttop; o->thl_top = thl_top; o->BCbotT = BCbotT; o->wtsurf = wtsurf;
  for (int k = 0; k <= K; k++) o->thlpcar[k] = thlpcar ? thlpcar[k] : 0.;
}
void orc_set_buoycorr(orc_t *o, int lbuoycorr, double Rigc) { o->lbuoycorr = lbuoycorr; o->Rigc = Rigc; }   /* NAMSUBGRID, src/modsubgriddata.f90:41,44 */
double *orc_thermo_profile(orc_t *o, const char *name) {
  if (!strcmp(name, "thl0av")) return o->thl0av;
  if (!strcmp(name, "thvh")) return o->thvh;
  return NULL;
}
/* avexy_ibm (src/modmpi.f90:623-664, lnan = .false.) with a real mask array (1 fluid): II = interior of mask, all ones without IBM */
static void avexy_mask(const orc_t *o, double *aver, const double *var, int is_tend, const double *mask, int w_kb_zero) {
  const int I = o->itot, J = o->jtot, K = o->ktot;
  long *cnt = (long *)calloc(K + 2, sizeof(long));
  double *s = (double *)calloc(K + 2, sizeof(double)), *sall = (double *)calloc(K + 2, sizeof(double));
  for (int k = 1; k <= K + 1; k++)
    for (int j = 1; j <= J; j++)
      for (int i = 1; i <= I; i++) {
        const double v = is_tend ? T(var, i, j, k) : F(var, i, j, k);
        int II = mask ? (F(mask, i, j, k) != 0.) : 1;
        if (mask && w_kb_zero && k == 1) II = 0;           /* IIw(:,:,kb) = 0 with IBM (src/modibm.f90:2172) */
        s[k] += v * II; sall[k] += v; cnt[k] += II;
      }
  for (int k = 1; k <= K + 1; k++) {
    long d = cnt[k];
    double a = s[k];
    if (k == 1 && d == 0) { a = sall[k]; d = cnt[K]; }
    aver[k - 1] = d == 0 ? -999. : a / (double)d;
  }
  free(cnt); free(s); free(sall);
}
void orc_thermodynamics(orc_t *o) {
  if (!o->ltempeq) return;
  const int I = o->itot, J = o->jtot, K = o->ktot;
  const double eps1 = 1.e-10;
  avexy_mask(o, o->thl0av, o->thl0, 0, o->libm ? o->mask[3] : NULL, 0);          /* diagfld :270 */
  for (int k = 1; k <= K + 1; k++)                                               /* calc_halflev :518-526 */
    for (int j = 1; j <= J; j++)
      for (int i = 1; i <= I; i++)
        F(o->thl0h, i, j, k) = (F(o->thl0, i, j, k) * M(dzf, k - 1) + F(o->thl0, i, j, k - 1) * M(dzf, k)) / (2 * M(dzh, k));
  for (int j = 1; j <= J; j++)
    for (int i = 1; i <= I; i++) F(o->thl0h, i, j, 1) = o->thls;
  /* calthv, dry branch :213-235 */
  memset(o->dthvdz, 0, nT(o) * sizeof(double));
  for (int k = 1; k <= K + 1; k++)
    for (int j = 1 - o->jh; j <= J + o->jh; j++)
      for (int i = 1 - o->ih; i <= I + o->ih; i++) T(o->thv0h, i, j, k) = F(o->thl0h, i, j, k);
  for (int k = 2; k <= K; k++)
    for (int j = 1; j <= J; j++)
      for (int i = 1; i <= I; i++)
        T(o->dthvdz, i, j, k) = (F(o->thl0, i, j, k + 1) - F(o->thl0, i, j, k - 1)) / (M(dzh, k + 1) + M(dzh, k));
  for (int k = 1; k <= K; k++)
    for (int j = 1; j <= J; j++)
      for (int i = 1; i <= I; i++)
        if (fabs(T(o->dthvdz, i, j, k)) < eps1) T(o->dthvdz, i, j, k) = copysign(eps1, T(o->dthvdz, i, j, k));
  avexy_mask(o, o->thvh, o->thv0h, 1, o->libm ? o->mask[2] : NULL, 1);           /* :76 */
  o->thvh[0] = o->thl0av[0];                                                     /* :87: th0av(kb) * (1 + 0 - 0), dry */
  if (fabs(o->thvh[1]) < eps1) o->thvh[1] = o->thl0av[1];                        /* :88-90 */
}

/* ---- masscorr, volume-flow branches --------------------------------------------------------------------------
 * src/modforces.f90:394-420 (u) and :470-495 (v) with avexy_ibm (src/modmpi.f90:623-664, lnan = .false.). */
void orc_set_masscorr(orc_t *o, int luvolflowr, int lvvolflowr, double uflowrate, double vflowrate, const int *IIu, const int *IIv) {
  const int I = o->itot, J = o->jtot, K = o->ktot;
  const size_t n = (size_t)I * J * (K + 1);
  o->luvolflowr = luvolflowr; o->lvvolflowr = lvvolflowr; o->uflowrate = uflowrate; o->vflowrate = vflowrate;
  for (int c = 0; c < 2; c++) {
    int **II = c ? &o->IIv : &o->IIu, **IIs = c ? &o->IIvs : &o->IIus;
    const int *src = c ? IIv : IIu;
    if (!*II) { *II = (int *)malloc(n * sizeof(int)); *IIs = (int *)malloc((K + 1) * sizeof(int)); }
    for (size_t q = 0; q < n; q++) (*II)[q] = src ? src[q] : 1;
    for (int k = 0; k <= K; k++) {
      int s = 0;
      for (size_t q = 0; q < (size_t)I * J; q++) s += (*II)[(size_t)k * I * J + q];
      (*IIs)[k] = s;
    }
  }
}
static void avexy_ibm(const orc_t *o, double *aver, const double *var /* tendency-shaped or F-shaped */, int is_tend, const int *II, const int *IIs) {
  const int I = o->itot, J = o->jtot, K = o->ktot;
  for (int k = 1; k <= K + 1; k++) {
    double s = 0., sall = 0.;
    for (int j = 1; j <= J; j++)
      for (int i = 1; i <= I; i++) {
        const double v = is_tend ? T(var, i, j, k) : F(var, i, j, k);
        s += v * II[(i - 1) + (size_t)I * ((j - 1) + (size_t)J * (k - 1))];
        sall += v;
      }
    int d = IIs[k - 1];
    if (k == 1 && d == 0) { s = sall; d = IIs[K - 1]; }   /* modmpi.f90:649-652 */
    aver[k - 1] = d == 0 ? -999. : s / d;
  }
}
void orc_masscorr(orc_t *o, double dt, int rk3step) {
  const int I = o->itot, J = o->jtot, K = o->ktot;
  const double rk3coef = dt / (4. - (double)rk3step), rk3coefi = 1 / rk3coef;
  double *vol = (double *)malloc(2 * (K + 1) * sizeof(double)), *volold = vol + K + 1;
  const double zhtop = M(zh, K + 1);
  for (int c = 0; c < 2; c++) {
    if (!(c ? o->lvvolflowr : o->luvolflowr)) continue;
    double *tp = c ? o->vp : o->up;
    avexy_ibm(o, vol, tp, 1, c ? o->IIv : o->IIu, c ? o->IIvs : o->IIus);
    avexy_ibm(o, volold, c ? o->vm : o->um, 0, c ? o->IIv : o->IIu, c ? o->IIvs : o->IIus);
    double s1 = 0., s2 = 0.;
    for (int k = 1; k <= K; k++) { s1 += vol[k - 1] * M(dzf, k); s2 += volold[k - 1] * M(dzf, k); }
    const double outflow = rk3coef * s1 / zhtop, flowrateold = s2 / zhtop;
    const double def = (c ? o->vflowrate : o->uflowrate) - (outflow + flowrateold);
    if (c) o->vdef = def; else o->udef = def;
    for (int k = 1; k <= K; k++)
      for (int j = 1; j <= J; j++)
        for (int i = 1; i <= I; i++) T(tp, i, j, k) = T(tp, i, j, k) + def * rk3coefi;
  }
  free(vol);
}
void orc_masscorr_get(orc_t *o, double *udef, double *vdef) { *udef = o->udef; *vdef = o->vdef; }

void orc_substep(orc_t *o, double *dt, int *rk3step, double dtmax, int ladaptive, double courant, double diffnr) {
  orc_tstep_update(o, dt, courant, diffnr, dtmax, ladaptive, rk3step, NULL, NULL);
  orc_advection(o);
  orc_subgrid(o);
  orc_bottom(o);         /* src/program.f90:152 */
  orc_forces(o);         /* src/program.f90:158 */
  orc_ibm_diffcorr(o);   /* the in-scope part of ibmwallfun, src/program.f90:166 */
  if (o->luvolflowr || o->lvvolflowr) orc_masscorr(o, *dt, *rk3step);   /* src/program.f90:169 */
  orc_ibmnorm(o);        /* src/program.f90:171 */
  orc_poisson(o, *dt, *rk3step);
  orc_tstep_integrate(o, *dt, *rk3step);
  orc_halos(o);
  orc_boundary(o);
  orc_thermodynamics(o); /* src/program.f90:212 */
}
