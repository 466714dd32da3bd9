/* udales_oracle.h — CPU restatement (parity ORACLE) of the uDALES dynamics hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (u-dales_b200/, include/)
 * may include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, and only as the checker
 * or as the timed CPU baseline.
 *
 * Every function cites the reference file:line (relative to /root/reference) whose
 * loop nest, operand order and boundary handling it restates.  The reference is
 * Fortran (1-based, column-major, halos); the macros below reproduce its index
 * space so the loop bodies can be compared line by line.
 *
 * Single pencil (nprocx = nprocy = 1): lateral periodicity is applied by the local
 * wrap routines, exactly what the reference does when a direction is unsplit
 * (src/modboundary.f90:95-107).  Decomposition invariance (1e-9, reference's own
 * tolerance) lets this serve as the oracle for multi-GPU runs too.
 */
#ifndef UDALES_ORACLE_H
#define UDALES_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_cfg {
  int itot, jtot, ktot;
  int nsv;            /* passive scalars (kappa advection, halo 2)              */
  int BCtopm;         /* 1 freeslip, 2 noslip   (src/modglobal.f90:140-142)     */
  int lles;           /* .true. when any SGS model is on (modsubgrid.f90:118)   */
  int lvreman, lsmagorinsky;
  int iadv_sv;        /* 7 kappa (forced by modglobal.f90:556-559), 2 cd2, 1 upw */
  double xlen, ylen;
  double numol, prandtlmoli, prandtli, c_vreman, cs;
  double Uinf, Vinf;
  const double *zf;   /* ktot cell-centre heights (column 1 of prof.inp)        */
} orc_cfg;

typedef struct orc orc_t;

orc_t *orc_create(const orc_cfg *cfg);
void orc_destroy(orc_t *o);

/* field access: returns pointer to the first element of the Fortran array (incl. halos)
 * and its three extents.  Names: u0 v0 w0 um vm wm up vp wp pres0 p ekm ekh rhs
 * pup pvp pwp sv0 svm svp (scalar arrays have a 4th extent nsv). */
double *orc_field(orc_t *o, const char *name, int dims[4]);
/* 1-D metric arrays: dzf dzh dzfi dzhi xrt yrt a b c ; returns pointer to Fortran index
 * lo (written to *lo) */
double *orc_metric(orc_t *o, const char *name, int *lo, int *n);

/* hot-path routines, named after the reference procedures */
void orc_advection(orc_t *o);                       /* modadvection.f90:36   */
void orc_closure(orc_t *o);                         /* modsubgrid.f90:159 (+closurebc) */
void orc_subgrid(orc_t *o);                         /* modsubgrid.f90:128   */
void orc_fillps(orc_t *o, double dt, int rk3step);  /* modpois.f90:911      */
void orc_poisson_solve(orc_t *o, double *pz);       /* modpois.f90:440-712 on an (itot,jtot,ktot) array */
void orc_tderive(orc_t *o);                         /* modpois.f90:1001     */
void orc_poisson(orc_t *o, double dt, int rk3step); /* modpois.f90:419      */
void orc_tstep_update(orc_t *o, double *dt, double courant, double diffnr, double dtmax,
                      int ladaptive, int *rk3step, double *courtot, double *diffnrtot); /* modtstep.f90:49 */
void orc_tstep_integrate(orc_t *o, double dt, int rk3step);  /* modtstep.f90:171 */
void orc_halos(orc_t *o);                           /* modboundary.f90:67   */
void orc_boundary(orc_t *o);                        /* modboundary.f90:115  */
void orc_chkdiv(orc_t *o, double *divmax, double *divtot, double *divrms); /* modchecksim.f90:161 */
void orc_randomize(orc_t *o, const char *name, int n4, double ampl, int ir); /* modstartup.f90:2367 (all k = 1..ktot) */
/* one pass of program.f90:132-207 restricted to the in-scope calls */
void orc_substep(orc_t *o, double *dt, int *rk3step, double dtmax, int ladaptive,
                 double courant, double diffnr);

/* stand-alone real FFT helpers (FFTW r2c/c2r conventions + the reference's packing) */
void orc_rfft_packed(int n, double *line, int inverse);  /* modpois.f90:478-490 / 669-679 on one line */

/* forces (src/modforces.f90:46, neutral branch): dpdxl, dpdyl = ktot+1 values (kb:ke+kh) */
void orc_set_forcing(orc_t *o, const double *dpdxl, const double *dpdyl);
void orc_forces(orc_t *o);

/* bottom -> wfmneutral case 91 (src/modibm.f90:1998, src/modwallfunctions.f90:307-349) and the zero-flux scalar
 * correction (:2077-2091); masscorr, volume-flow branches (src/modforces.f90:394-420, 470-495).  IIu / IIv:
 * (itot, jtot, ktot+1) ints, 1 = fluid, or NULL (no blocks). */
void orc_set_bottom(orc_t *o, int lbottom, int BCbotm, int BCbots, double z0, double fkar);
void orc_bottom(orc_t *o);
/* BCbotm = 2 / BCbotT = 2: wfuno cases 91 / 92 with the stability functions unom / unoh (src/modwallfunctions.f90:24-260).
 * tcell: the uniform thl0(kb) of a run without temperature equation (Tcell of case 91 then). */
void orc_set_wfuno(orc_t *o, double z0h, double prandtlturb, double grav, double thls, double tcell);
double *orc_momfluxb(orc_t *o);
void orc_set_masscorr(orc_t *o, int luvolflowr, int lvvolflowr, double uflowrate, double vflowrate, const int *IIu, const int *IIv);
void orc_masscorr(orc_t *o, double dt, int rk3step);
void orc_masscorr_get(orc_t *o, double *udef, double *vdef);

/* temperature, dry: ltempeq with iadv_thl = cd2 — advecc_2nd + diffc on thl0 (momentum halo), fixed-flux
 * bottom (BCbotT = 1, src/modibm.f90:2033-2046), top flux / value (BCtopT = 1 / 2, src/modboundary.f90:208-221), buoyancy
 * and radiative tendency in forces (src/modforces.f90:70-109), thermodynamics (src/modthermodynamics.f90:55-121, dry).
 * Fields thl0 thlm thlp thl0h thv0h dthvdz through orc_field; profiles thl0av / thvh (ktot+1 values) through
 * orc_thermo_profile.  thlpcar: ktot+1 values or NULL. */
void orc_set_thermo(orc_t *o, int lbuoyancy, double grav, double thls, int BCtopT, double wttop, double thl_top,
                    int BCbotT, double wtsurf, const double *thlpcar);
void orc_thermodynamics(orc_t *o);
/* lbuoycorr / Rigc (NAMSUBGRID): buoyancy correction of the Vreman eddy viscosity, src/modsubgrid.f90:332-354; reads the
 * dthvdz of the last orc_thermodynamics */
void orc_set_buoycorr(orc_t *o, int lbuoycorr, double Rigc);
double *orc_thermo_profile(orc_t *o, const char *name);

/* immersed boundary masking: kind 0-3 = solid_u,v,w,c ; 4-7 = fluid-boundary points u,v,w,c;
 * ijk = n local 1-based (i,j,k) triples, point-major */
void orc_ibm_set_points(orc_t *o, int kind, int n, const int *ijk);
void orc_ibm_build_masks(orc_t *o);                 /* modibm.f90:153-192 */
double *orc_ibm_mask(orc_t *o, int m);              /* mask_u, mask_v, mask_w, mask_c */
void orc_ibmnorm(orc_t *o);                         /* modibm.f90:697 */
void orc_ibm_diffcorr(orc_t *o);                    /* modibm.f90:990-1164 (called from ibmwallfun :1211-1241) */

#ifdef __cplusplus
}
#endif
#endif
