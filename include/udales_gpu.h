/* udales_gpu.h — C-ABI of the B200-native uDALES dynamics core (libudales_gpu.so).
 *
 * This is the drop-in boundary for the per-RK3-substep hot path of uDALES
 * (reference = /root/reference @ 2fe4df1; file:line citations are relative to it).
 * The reference has no plugin API: the boundary is the set of argument-less Fortran
 * module procedures called from src/program.f90:134-207 which talk through public
 * module arrays (src/modfields.f90) and namelist switches (src/modglobal.f90).
 * Each entry point below replaces one of those procedures; the ISO_C_BINDING shim a
 * maintainer adds on the Fortran side is u-dales_b200/fortran/modgpu.f90 and the
 * patch points are listed in INTEGRATION.md.
 *
 * Conventions
 *  - plain C: ints, doubles, pointers, sizes.  No C++/torch types.
 *  - every function returns 0 on success, a negative UDGPU_E* code on failure and
 *    never aborts; udgpu_last_error() gives the text.  The Fortran shim turns a
 *    non-zero code into `write(0,*) ...; stop 1`, the reference's error convention
 *    (e.g. src/modpois.f90:896-898).
 *  - all 3-D arrays are Fortran column-major fp64 with the reference's halo shapes
 *    (src/modfields.f90:440-474): the pointer addresses element (ib-ih, jb-jh, klo).
 *  - one host thread per GPU/rank; calls are collective across ranks and ordered,
 *    like the MPI ranks of the reference.
 *  - the library owns all device memory and streams; host arrays stay owned by the
 *    caller and are touched only inside udgpu_push / udgpu_pull.
 *  - there is NO CPU fallback: without a CUDA device udgpu_init fails with
 *    UDGPU_ENODEV.
 */
#ifndef UDALES_GPU_H
#define UDALES_GPU_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UDGPU_ABI_VERSION 2

/* error codes */
#define UDGPU_OK 0
#define UDGPU_EINVAL (-1)   /* bad argument / unsupported switch combination          */
#define UDGPU_ENODEV (-2)   /* no CUDA device / wrong architecture                    */
#define UDGPU_ECUDA (-3)    /* CUDA runtime or driver error                           */
#define UDGPU_ENCCL (-4)    /* NCCL error                                             */
#define UDGPU_ESTATE (-5)   /* call order violated (e.g. poisson before init)         */
#define UDGPU_ENOMEM (-6)

/* field ids for push / pull / device_ptr (names = src/modfields.f90, src/modpois.f90:41,
 * src/modsubgriddata.f90) */
enum udgpu_field {
  UDGPU_U0 = 0, UDGPU_V0, UDGPU_W0,          /* (ib-ih:ie+ih, jb-jh:je+jh, kb-kh:ke+kh) */
  UDGPU_UM, UDGPU_VM, UDGPU_WM,              /* same                                    */
  UDGPU_UP, UDGPU_VP, UDGPU_WP,              /* (.., .., kb:ke+kh)  tendencies          */
  UDGPU_PRES0,                               /* as u0                                   */
  UDGPU_P,                                   /* as u0   (src/modpois.f90:83)            */
  UDGPU_EKM, UDGPU_EKH,                      /* as u0   (src/modsubgrid.f90:54-55)      */
  UDGPU_RHS,                                 /* (imax,jmax,ktot) no halo (modpois.f90:84) */
  UDGPU_SV0, UDGPU_SVM,                      /* (ib-ihc:.., jb-jhc:.., kb-khc:ke+khc, nsv) */
  UDGPU_SVP,                                 /* (.., .., kb:ke+khc, nsv)                */
  UDGPU_MOMFLUXB,                            /* as u0; allocated by udgpu_set_bottom (src/modfields.f90 momfluxb)  */
  UDGPU_THL0, UDGPU_THLM,                    /* as u0 (alloc_z, src/modfields.f90:495-497); with ltempeq only      */
  UDGPU_THLP,                                /* as up (src/modfields.f90:453); with ltempeq only                   */
  UDGPU_NFIELDS
};

/* Everything the hot path reads from modglobal / modsubgriddata / decomp_2d at init.
 * Filled by the Fortran shim right after `call initpois` (src/program.f90:89). */
typedef struct udgpu_cfg {
  int abi_version;               /* = UDGPU_ABI_VERSION                                  */
  int itot, jtot, ktot;          /* global grid            (namelist DOMAIN)             */
  int imax, jmax, kmax;          /* local z-pencil = zsize (src/modglobal.f90:613-636)   */
  int ih, jh, kh;                /* momentum halo          (src/modglobal.f90:586-599)   */
  int ihc, jhc, khc;             /* scalar halo            (src/modglobal.f90:602-609)   */
  int nsv;                       /* passive scalars        (namelist SCALARS)            */
  int zstart[3];                 /* 1-based global index of local (1,1,1) (decomp_2d zstart) */
  int nprocx, nprocy;            /* p_row, p_col           (src/modstartup.f90:676)      */
  int myidx, myidy;              /* pencil coordinates     (src/modmpi.f90)              */
  int BCxm, BCym;                /* 1 periodic             (src/modglobal.f90:97,121)    */
  int BCtopm;                    /* 1 freeslip 2 noslip    (src/modglobal.f90:140-142)   */
  int BCzp;                      /* 1 = tridiagonal solve in z (src/modpois.f90:146)     */
  int ipoiss;                    /* 0 = POISS_FFT2D        (src/modglobal.f90:389)       */
  int iadv_mom;                  /* 2 = cd2                (src/modglobal.f90:398)       */
  int iadv_sv;                   /* 7 kappa, 2 cd2         (src/modglobal.f90:397-399)   */
  int lles, lvreman, lsmagorinsky, loneeqn;  /* src/modsubgriddata.f90:39-42            */
  int ltempeq, lmoist;           /* ltempeq: dry temperature equation (thl0, thlm, thlp resident; udgpu_set_thermo);
                                    lmoist must be 0 (out of scope)       */
  double dx, dy;                 /* src/modglobal.f90:710-711                            */
  const double *dzf;             /* dzf(kb-kh:ke+kh): ktot+2 values (src/modglobal.f90:751-755) */
  const double *dzh;             /* dzh(kb:ke+kh):   ktot+1 values (src/modglobal.f90:757-760)  */
  const double *delta;           /* delta(kb:ke+kh) at any i (x uniform) ktot+1 values (:793-797); may be NULL (computed) */
  double numol, prandtlmoli, prandtli;   /* src/modglobal.f90:300-303, modsubgrid.f90:117 */
  double c_vreman, cs;           /* src/modsubgriddata.f90:55,61 ; cs = -1 -> (cm^3/ceps)^1/4 */
  double Uinf, Vinf;             /* noslip top velocity    (src/modglobal.f90)           */
  double e12min;
  int device;                    /* CUDA device ordinal, -1 = LOCAL_RANK / current        */
  int flags;                     /* UDGPU_F_* */
  int iadv_thl;                  /* 2 = cd2 (0 = unset -> iadv_mom, src/modglobal.f90:549); ABI 2 */
} udgpu_cfg;

#define UDGPU_F_NO_LAZY_FUSION 1  /* run every call eagerly as its own kernel(s) (debug / parity bisecting) */
#define UDGPU_F_NO_GRAPH 2        /* do not capture the substep into a CUDA graph                          */
#define UDGPU_F_NCCL_TRANSPOSE 4  /* multi-GPU: ncclSend/Recv all-to-all instead of the peer-store (NVLink P2P) fused transposes */
#define UDGPU_F_V1_KERNELS 8      /* use the direct one-thread-per-cell kernels instead of the TMA-staged ones (cross-check) */
#define UDGPU_F_NO_HALO_FUSION 16 /* closure / tderive+integrate do not write their own halo and ghost cells: separate wrap,
                                     closurebc, bcp, halos and boundary kernels run instead (cross-check)                    */

typedef struct udgpu udgpu_t;     /* opaque */

/* ---- life cycle -------------------------------------------------------------------- */
/* replaces: allocation part of initfields/initpois/initsubgrid on the device
 * (src/modfields.f90:422, src/modpois.f90:66, src/modsubgrid.f90:44).
 * nccl_uid: 128-byte ncclUniqueId obtained from udgpu_nccl_unique_id on rank 0 and broadcast by
 * the host (MPI_Bcast in Fortran, torch.distributed in the tests); NULL when nprocx*nprocy == 1. */
int udgpu_nccl_unique_id(void *uid128);
int udgpu_init(const udgpu_cfg *cfg, const void *nccl_uid, udgpu_t **out);
int udgpu_finalize(udgpu_t *h);
const char *udgpu_last_error(void);
int udgpu_abi_version(void);

/* ---- residency control ------------------------------------------------------------- */
/* host <-> device copies of one field in the reference's own array shape.  n4 selects the
 * scalar index for SV0/SVM/SVP (0-based), ignored otherwise.  Asynchronous on the library
 * stream when host memory is pinned; udgpu_sync waits. */
int udgpu_push(udgpu_t *h, int field, int n4, const double *host);
int udgpu_pull(udgpu_t *h, int field, int n4, double *host);
int udgpu_field_count(udgpu_t *h, int field, size_t *count, int dims[3]);
int udgpu_device_ptr(udgpu_t *h, int field, int n4, void **dptr);   /* zero-copy interop */
/* sparse residency for host add-ons that touch few cells per substep (the facet wall functions of ibmwallfun,
 * src/modibm.f90:1286-1860: they read u0 v0 w0 thl0 next to the walls and add to up vp wp thlp at the fluid-boundary
 * points).  offsets: n 0-based linear offsets into the field's Fortran array (halos included, as udgpu_field_count
 * describes it).  pull: out[q] = field(offsets[q]).  add: field(offsets[q]) += vals[q], tendencies only; a point may occur
 * several times.  Both synchronise; both leave a lazily pending forces / masscorr pending. */
int udgpu_pull_points(udgpu_t *h, int field, int n4, long long n, const long long *offsets, double *out);
int udgpu_add_points(udgpu_t *h, int field, int n4, long long n, const long long *offsets, const double *vals);
int udgpu_sync(udgpu_t *h);
int udgpu_host_register(void *ptr, size_t bytes);    /* pin a Fortran array once (cudaHostRegister) */
int udgpu_host_unregister(void *ptr);

/* ---- the hot path: one entry point per reference procedure ------------------------- */
/* src/modtstep.f90:49  tstep_update.  In/out: dt, rk3step.  Also returns the two global maxima. */
int udgpu_tstep_update(udgpu_t *h, double *dt, double courant, double diffnr, double dtmax,
                       int ladaptive, int *rk3step, double *courtot, double *diffnrtot);
/* src/modadvection.f90:36  advection (advecu/v/w_2nd + per-scalar advecc_kappa / advecc_2nd) */
int udgpu_advection(udgpu_t *h);
/* src/modsubgrid.f90:128  subgrid (closure + closurebc + diffu/v/w + diffc) */
int udgpu_subgrid(udgpu_t *h);
/* src/modsubgrid.f90:159  closure only (ekm, ekh incl. ghost cells) */
int udgpu_closure(udgpu_t *h);
/* src/modpois.f90:419  poisson = fillps + FFT2D solve + tderive */
int udgpu_poisson(udgpu_t *h, double dt, int rk3step);
/* src/modpois.f90:440-712 core only: rhs (imax,jmax,ktot) z-pencil in, p same shape out (host
 * pointers; rhs == p allowed).  The "Poisson solves/s" unit of BASELINE.json. */
int udgpu_poisson_solve(udgpu_t *h, const double *rhs, double *p);
/* same on the resident UDGPU_RHS buffer, no host traffic */
int udgpu_poisson_solve_resident(udgpu_t *h);
/* src/modpois.f90:911 fillps (+bcpup) -> UDGPU_RHS and p interior ; src/modpois.f90:1001 tderive (+bcp) */
int udgpu_fillps(udgpu_t *h, double dt, int rk3step);
int udgpu_tderive(udgpu_t *h);
/* src/modtstep.f90:171  tstep_integrate */
int udgpu_tstep_integrate(udgpu_t *h, double dt, int rk3step);
/* src/modboundary.f90:67  halos ; src/modboundary.f90:115  boundary (periodic / freeslip / noslip subset) */
int udgpu_halos(udgpu_t *h);
int udgpu_boundary(udgpu_t *h);
/* src/modchecksim.f90:161  chkdiv: max |div|, sum div*dV, and RMS(div) (parity metric) */
int udgpu_divergence(udgpu_t *h, double *divmax, double *divtot, double *divrms);
/* one pass of src/program.f90:132-212 restricted to the calls of this header, in the reference's order:
 * tstep_update, advection, subgrid, [bottom, forces, ibm_diffcorr, masscorr, ibmnorm,] poisson, tstep_integrate, halos,
 * boundary [, thermodynamics]. */
int udgpu_substep(udgpu_t *h, double *dt, int *rk3step, double dtmax, int ladaptive,
                  double courant, double diffnr);

/* ---- resident-channel glue (next tier) ------------------------------------------------------ */
/* src/modforces.f90:46 forces, neutral branch: up -= dpdxl(k), vp -= dpdyl(k), wp(kb) = 0.  udgpu_set_forcing copies the
 * two profiles (dpdxl(kb:ke+kh), dpdyl(kb:ke+kh): ktot+1 values each, src/modfields.f90) to the device; udgpu_forces is
 * the call of src/program.f90:158 (applied inside the fused tderive+integrate kernel unless something looks at the
 * tendencies first).  udgpu_substep calls it after subgrid once a forcing has been set. */
int udgpu_set_forcing(udgpu_t *h, const double *dpdxl, const double *dpdyl);
int udgpu_forces(udgpu_t *h);

/* src/modibm.f90:1998 bottom with lbottom = .true.; BCbotm = 3: wfmneutral(.., 91) (src/modwallfunctions.f90:307-349) on
 * up, vp (k = kb) and momfluxb, plus the zero-flux scalar bottom correction (BCbots = 1, src/modibm.f90:2077-2091).
 * udgpu_set_bottom takes the namelist values (z0, fkar = von Karman constant); udgpu_bottom is the call of
 * src/program.f90:152; udgpu_substep calls it after subgrid once lbottom is set. */
int udgpu_set_bottom(udgpu_t *h, int lbottom, int BCbotm, int BCbots, double z0, double fkar);
/* BCbotm = 2 (the namelist default) / BCbotT = 2: wfuno cases 91 / 92 with the stability functions unom / unoh
 * (src/modwallfunctions.f90:24-260).  z0h, prandtlturb (src/modglobal.f90:304), grav, thls = wall temperature; tcell = the
 * uniform thl0(kb) of a run WITHOUT temperature equation (the reference evaluates the stability correction with
 * thl0 = thlprof and thls even then); with ltempeq the resident thl0 is used. */
int udgpu_set_wfuno(udgpu_t *h, double z0h, double prandtlturb, double grav, double thls, double tcell);
int udgpu_bottom(udgpu_t *h);
/* src/modforces.f90:328 masscorr, volume-flow branches (luvolflowr / lvvolflowr, :394-420 / :470-495): masked slab means of
 * the tendency and of um / vm (avexy_ibm, src/modmpi.f90:623-664; summed over all ranks), def = flowrate - (rk3coef <up> + <um>),
 * up += def / rk3coef.  The masks IIu / IIv are the interiors of mask_u / mask_v of udgpu_ibm_commit (all ones without
 * IBM).  udef / vdef may be NULL (then the call does not synchronise the host).  udgpu_substep calls it between the
 * IBM diffusion corrections and ibmnorm (src/program.f90:169) once a flow rate has been set. */
int udgpu_set_masscorr(udgpu_t *h, int luvolflowr, int lvvolflowr, double uflowrate, double vflowrate);
int udgpu_masscorr(udgpu_t *h, double dt, int rk3step, double *udef, double *vdef);

/* ---- temperature, dry (next tier; needs cfg.ltempeq = 1, iadv_thl = cd2) ----------------------------
 * With ltempeq the library keeps thl0, thlm, thlp resident and the existing entry points do the temperature part of the
 * reference procedures they replace: advection -> advecc_2nd(thl0, thlp) (src/modadvection.f90:67-69); subgrid ->
 * diffc(thl0, thlp) (src/modsubgrid.f90:146) and closurebc's fluxtop on thl (src/modboundary.f90:417-420); bottom ->
 * fixed-flux temperature branch (BCbotT = 1, src/modibm.f90:2033-2046); forces -> buoyancy on wp and thlp += thlpcar
 * (src/modforces.f90:70-83, 103-109); ibm_diffcorr -> diffc_corr(thl0, thlp) (src/modibm.f90:1225); ibmnorm -> solid(..,
 * thlm, thlp, <thl0av>, mask_c) + advecc2nd_corr_liberal (:714-722); tstep_integrate (src/modtstep.f90:244,325,334);
 * halos (xT_periodic / yT_periodic, src/modboundary.f90:541-556) and boundary (BCtopT = 1 fluxtop(.., ekh, wttop) / 2
 * valuetop(.., thl_top), :208-221).
 * udgpu_set_thermo takes the namelist values (PHYSICS: lbuoyancy; BC: BCtopT, BCbotT, wttop, thl_top, wtsurf; thls) and the
 * radiative tendency profile thlpcar(kb:ke+kh) (ktot+1 values, may be NULL = 0).  BCbotT = 2 needs udgpu_set_wfuno; the
 * facet heat fluxes of wallfunheat stay with the host.
 * udgpu_thermodynamics is the call of src/program.f90:212 (src/modthermodynamics.f90:55-121, lmoist = .false.): slab means
 * thl0av (mask IIc) and thvh (thv0h = thl0h, mask IIw, kb / kb+1 overrides), summed over all ranks; the hydrostatic
 * pressure / exner / density profiles (fromztop) are not on the dry path and are not computed.  It must follow every
 * change of thl0 (as in the reference) and run once before the first forces(); udgpu_substep calls it last.
 * udgpu_thermo_profile copies a profile to the host: which = 0 thl0av(kb:ke+kh), 1 thvh(kb:ke+kh) (ktot+1 values). */
int udgpu_set_thermo(udgpu_t *h, int lbuoyancy, double grav, double thls, int BCtopT, double wttop, double thl_top,
                     int BCbotT, double wtsurf, const double *thlpcar);
int udgpu_thermodynamics(udgpu_t *h);
/* NAMSUBGRID lbuoycorr / Rigc: buoyancy correction of the Vreman eddy viscosity for stable stratification
 * (src/modsubgrid.f90:332-354), active with lbuoyancy; applied by udgpu_closure / udgpu_subgrid */
int udgpu_set_buoycorr(udgpu_t *h, int lbuoycorr, double Rigc);
int udgpu_thermo_profile(udgpu_t *h, int which, double *host);

/* ---- immersed-boundary masking (next tier; src/modibm.f90) ------------------------------- */
/* point lists of modibm: kind 0-3 = solid_info_{u,v,w,c}%solpts_loc, 4-7 = bound_info_{u,v,w,c}%bndpts_loc; n points,
 * local 1-based (i,j,k).  layout 0: point-major triples [i0,j0,k0,i1,...]; layout 1: the Fortran array (n,3) as it
 * lies in memory (all i, then all j, then all k). */
enum udgpu_ibm_kind { UDGPU_IBM_SOLID_U = 0, UDGPU_IBM_SOLID_V, UDGPU_IBM_SOLID_W, UDGPU_IBM_SOLID_C,
                      UDGPU_IBM_BOUND_U, UDGPU_IBM_BOUND_V, UDGPU_IBM_BOUND_W, UDGPU_IBM_BOUND_C };
int udgpu_ibm_set_points(udgpu_t *h, int kind, int n, const int *ijk, int layout);
/* builds mask_u, mask_v, mask_w, mask_c on the device as initibm does (src/modibm.f90:153-192: 1, ground level 0,
 * solid points 0, halo exchange) and switches the IBM calls of udgpu_substep on.  m = 0..3 can be pulled for checking. */
int udgpu_ibm_commit(udgpu_t *h);
int udgpu_ibm_pull_mask(udgpu_t *h, int m, double *host);
/* src/modibm.f90:697 ibmnorm (momentum + scalars) */
int udgpu_ibmnorm(udgpu_t *h);
/* src/modibm.f90:1211-1213,1240-1242: diffu_corr, diffv_corr, diffw_corr, diffc_corr (the part of ibmwallfun on the
 * resident path; the wall-function stresses themselves stay with the host) */
int udgpu_ibm_diffcorr(udgpu_t *h);

/* One full RK3 time step on HOST arrays (the literal drop-in for a host-resident model): pushes u0,v0,w0,pres0
 * (reference shapes; um=u0 at the start of a time step, src/modtstep.f90:330-338), runs the three substeps of
 * src/program.f90:132-207 on the device and pulls the same four arrays back.  Pin the arrays once with
 * udgpu_host_register for full PCIe bandwidth.  dt is in/out as in udgpu_tstep_update. */
int udgpu_rk3_step_host(udgpu_t *h, double *u0, double *v0, double *w0, double *pres0, double *dt,
                        double dtmax, int ladaptive, double courant, double diffnr);

/* ---- measurement hooks (used by bench.py; no effect on results) --------------------- */
/* device time in ms of the most recent launch group of one hot-path kernel family, measured
 * with CUDA events on the library stream.  which: 0 mom_tend, 1 closure, 2 poisson core,
 * 3 fillps, 4 tderive+integrate, 5 halos+boundary. */
int udgpu_profile_enable(udgpu_t *h, int on);
int udgpu_profile_get(udgpu_t *h, int which, double *ms_total, long *launches);
int udgpu_profile_reset(udgpu_t *h);
long udgpu_launch_count(udgpu_t *h);          /* kernels launched by this library so far */
int udgpu_stream(udgpu_t *h, void **cuda_stream);
/* udgpu_profile_enable(h, 2) records a CUDA event at every stage of the substep on the stream it runs on (main, barrier,
 * copy streams); udgpu_trace_dump waits for the device and writes one line per mark: "t_ms lane chunk label". */
int udgpu_trace_dump(udgpu_t *h, const char *path);

#ifdef __cplusplus
}
#endif
#endif /* UDALES_GPU_H */
