/* ctsm_b200.h — C ABI of the B200-native CTSM biogeophysics hot path.
 *
 * Every entry point below is what a Fortran `bind(C)` interface in the host
 * model binds (see INTEGRATION.md for the shim module).  Each one replaces one
 * reference subroutine, cited as file:line relative to the CTSM checkout:
 *
 *   ctsm_b200_tridiagonal        src/biogeophys/TridiagonalMod.F90:23
 *   ctsm_b200_banddiagonal       src/biogeophys/BandDiagonalMod.F90:29   (LAPACK dgbsv semantics)
 *   ctsm_b200_dgtsv_batch        src/biogeophys/SoilWaterMovementMod.F90:1279-1299 (LAPACK dgtsv call site)
 *   ctsm_b200_soilwater          src/biogeophys/SoilWaterMovementMod.F90:240  (moisture_form, :976)
 *   ctsm_b200_soiltemperature    src/biogeophys/SoilTemperatureMod.F90:92
 *   ctsm_b200_canopyfluxes       src/biogeophys/CanopyFluxesMod.F90:191  (+ PhotosynthesisMod.F90:2704 PHS)
 *   ctsm_b200_set_filters        src/main/filterMod.F90:303   (setFiltersOneGroup)
 *   ctsm_b200_set_exposedvegp_filter  src/main/filterMod.F90:595
 *   ctsm_b200_begin_water_column_balance  src/biogeophys/BalanceCheckMod.F90:171  (+ TotalWaterAndHeatMod.F90:92)
 *   ctsm_b200_water_gridcell_balance      src/biogeophys/BalanceCheckMod.F90:132  (+ TotalWaterAndHeatMod.F90:92,144)
 *   ctsm_b200_balancecheck       src/biogeophys/BalanceCheckMod.F90:445,859
 *   ctsm_b200_soilfluxes         src/biogeophys/SoilFluxesMod.F90:37   (+ p2c, src/main/subgridAveMod.F90:292)
 *   ctsm_b200_patch2col          src/main/clm_driver.F90:1655          (clm_drv_patch2col)
 *
 * Conventions:
 *   - plain pointers and sizes only; no C++/torch types cross this boundary;
 *   - all subgrid indices and filters are 1-based proc-local, exactly as the
 *     Fortran passes them; arrays are column-major, subgrid index fastest;
 *   - `mem` says where the pointers live: CTSM_MEM_DEVICE (device-resident
 *     state, kernels launched on the context stream, asynchronous) or
 *     CTSM_MEM_HOST (host arrays as the Fortran owns them: the library stages
 *     them through its device mirrors, synchronous);
 *   - return value 0 = success; nonzero = the reference would have called
 *     endrun(); ctsm_status_t then carries the subgrid level/index and the
 *     message text of the reference's abort site (abortutils.F90:68-97);
 *   - there is no CPU fallback: every compute entry point returns
 *     CTSM_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef CTSM_B200_H
#define CTSM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTSM_B200_ABI_VERSION 5

/* fixed vertical structure the kernels are compiled for (clm_varpar.F90:43-54,
 * 290-292; namelist_defaults_ctsm.xml:254,511).  ctsm_b200_init refuses any
 * other configuration. */
#define CTSM_NLEVSNO 12
#define CTSM_NLEVGRND 25
#define CTSM_NLEVSOI 20
#define CTSM_NVEGWCS 4
#define CTSM_NLEVCAN 1
#define CTSM_NLEVLAK 10
#define CTSM_MXPFT 78

/* decompMod.F90:60-68 */
typedef struct ctsm_bounds_t {
  int32_t begg, endg;
  int32_t begl, endl;
  int32_t begc, endc;
  int32_t begp, endp;
  int32_t begCohort, endCohort;
  int32_t level;
  int32_t clump_index;
} ctsm_bounds_t;

/* decompMod.F90:26-32 */
enum { CTSM_SUBGRID_UNSPECIFIED = -1, CTSM_SUBGRID_LNDGRID = 0, CTSM_SUBGRID_GRIDCELL = 1,
       CTSM_SUBGRID_LANDUNIT = 2, CTSM_SUBGRID_COLUMN = 3, CTSM_SUBGRID_PATCH = 4 };

/* landunit_varcon.F90:19-30 */
enum { CTSM_ISTSOIL = 1, CTSM_ISTCROP = 2, CTSM_ISTICE = 4, CTSM_ISTDLAK = 5, CTSM_ISTWET = 6,
       CTSM_ISTURB_MIN = 7, CTSM_ISTURB_MAX = 9 };

enum { CTSM_MEM_DEVICE = 0, CTSM_MEM_HOST = 1,
       /* host arrays, but OUT fields are not uploaded first: elements outside
        * the filter / active levels come back undefined */
       CTSM_MEM_HOST_NOPRESERVE = 3 };

enum {
  CTSM_OK = 0,
  CTSM_ERR_NO_DEVICE = 1,       /* no usable CUDA device (there is no CPU fallback) */
  CTSM_ERR_BAD_ARG = 2,         /* NULL pointer, bad bounds, unsupported config  */
  CTSM_ERR_NOMEM = 3,           /* device / pinned-host allocation failed (cudaErrorMemoryAllocation) */
  CTSM_ERR_CUDA = 4,            /* any other CUDA runtime error; ctsm_b200_last_cuda_error() names it */
  CTSM_ERR_DGBSV = 10,          /* BandDiagonalMod.F90:200-213  "BandDiagonal ERROR: dgbsv returned error code" */
  CTSM_ERR_DGTSV = 11,          /* SoilWaterMovementMod.F90:1295 "soilwater_moisture_form:: problem with the lapack solver" */
  CTSM_ERR_FORC_HGT = 12,       /* CanopyFluxesMod.F90:997-1002  forcing height below canopy height */
  CTSM_ERR_GS_NEG = 13,         /* PhotosynthesisMod.F90:3686-3690 negative stomatal conductance */
  CTSM_ERR_BRENT = 14,          /* PhotosynthesisMod.F90:4134-4137 root must be bracketed for brent */
  CTSM_ERR_QUADRATIC = 15,      /* quadraticMod.F90:42-58 */
  CTSM_ERR_URBAN = 16,          /* urban column in filter: outside the hot path */
  CTSM_ERR_SNOW_NEGATIVE = 18,  /* SnowHydrologyMod.F90:1244-1287 UpdateState_TopLayerFluxes: top-layer ice / liquid significantly negative */
  CTSM_ERR_RH = 17,             /* HumanIndexMod.F90:1016-1022 Wet_BulbS: 2 m relative humidity outside [0, 100] */
  CTSM_ERR_BALANCE = 20         /* BalanceCheckMod.F90:640-659,1060-1114 thresholds exceeded */
};

typedef struct ctsm_status_t {
  int32_t code;           /* CTSM_OK or CTSM_ERR_* of the FIRST (lowest-index) failing point */
  int32_t subgrid_level;  /* CTSM_SUBGRID_* as passed to endrun(subgrid_level=) */
  int32_t subgrid_index;  /* 1-based proc-local index as passed to endrun(subgrid_index=) */
  int32_t info;           /* LAPACK info / iteration count / site-specific integer */
  double  value;          /* offending value where the reference prints one */
  int32_t n_warnings;     /* count of points where the reference only write(iulog)s */
  int32_t reserved;
  char    msg[160];       /* message text of the reference's endrun call */
} ctsm_status_t;

/* Scalars the reference reads from namelists / parameter files at init.
 * Defaults (ctsm_b200_default_params) follow namelist_defaults_ctsm.xml for
 * clm6_0 physics (SURVEY.md section 5 "Config / flag system"). */
typedef struct ctsm_params_t {
  int32_t abi_version;
  int32_t device;                 /* CUDA device ordinal */
  int32_t nlevsno, nlevgrnd, nlevsoi;
  double  dtime;                  /* get_step_size_real() */
  /* soilwater_movement_inparm, SoilWaterMovementMod.F90:104-236 */
  int32_t upper_boundary_condition;  /* 1 = bc_flux (only supported value)        */
  int32_t lower_boundary_condition;  /* 2 = bc_zero_flux, 1 = bc_flux             */
  int32_t flux_calculation;          /* 1 = inexpensive, 42 = expensive           */
  double  dtmin, verySmall, xTolerUpper, xTolerLower;
  double  e_ice;                     /* params_inst%e_ice, SoilWaterMovementMod.F90:34 */
  /* SoilTemperatureMod.F90:753-775 */
  int32_t snow_thermal_cond_method;      /* 1 = Jordan1991, 2 = Sturm1997 */
  int32_t snow_thermal_cond_glc_method;  /* 1 = Jordan1991, 2 = Sturm1997 */
  /* canopyfluxes_inparm + CanopyFluxesMod params_inst (CanopyFluxesMod.F90:61-70,115-117,174-186) */
  int32_t itmax_canopy_fluxes;          /* 40 */
  int32_t use_undercanopy_stability;    /* 0 */
  int32_t use_biomass_heat_storage;     /* 1 (clm6_0) */
  int32_t z0param_method;               /* 1 = ZengWang2007, 2 = Meier2022 (clm6_0) */
  int32_t soil_resis_method;            /* 0 = Lee-Pielke beta (do_soilevap_beta), 1 = SL14 (do_soil_resistance_sl14) */
  int32_t use_hydrstress;               /* 1 (only supported value) */
  int32_t use_luna;                     /* 1: vcmx25_z/jmx25_z inputs drive C3 non-crop vcmax (PhotosynthesisMod.F90:3353,3395) */
  int32_t stomatalcond_mtd;             /* 1 = Ball-Berry1987, 2 = Medlyn2011 */
  int32_t light_inhibit;                /* 1 */
  int32_t modifyphoto_and_lmr_forcrop;  /* 1 */
  double  lai_dl, z_dl, a_coef, a_exp, csoilc, cv, wind_min;
  double  zetamaxstable;                /* frictionvel_inst%zetamaxstable: 0.5 or 2.0 */
  double  leaf_mr_vcm;                  /* canopystate_inst%leaf_mr_vcm = 0.015 */
  /* photo_params_type scalars, PhotosynthesisMod.F90:93-118 */
  double  act25, fnr, cp25_yr2000, kc25_coef, ko25_coef, fnps, theta_psii, theta_ip;
  double  vcmaxha, jmaxha, tpuha, lmrha, kcha, koha, cpha;
  double  vcmaxhd, jmaxhd, tpuhd, lmrhd, lmrse;
  double  tpu25ratio, kp25ratio, vcmaxse_sf, jmaxse_sf, tpuse_sf, jmax25top_sf;
  /* BalanceCheckMod.F90:74-95 */
  int32_t balance_skip_steps;           /* set by ctsm_b200_balancecheck_init */
  int32_t npft_table;                   /* length of every PFT parameter table: (CTSM_MXPFT+1) x number of parameter-set
                                         * members; patch%itype indexes the table directly, so a perturbed-parameter
                                         * ensemble (BASELINE config 5) runs member m's patches with itype = m*(mxpft+1)+pft.
                                         * Default CTSM_MXPFT+1 (one parameter set: pftcon as the reference holds it). */
  int32_t calc_human_stress_indices;    /* 0 = NONE, 1 = FAST (clm5/clm6 default; HumanIndexMod.F90:496-547), ALL is not built */
  /* BiogeophysPreFluxCalcs: FrictionVelocityMod.F90 (zlnd, zsno, zglc: parameter-file scalars), clm_varctl use_z0m_snowmelt,
   * SurfaceResistanceMod.F90:35-36 (d_max, frac_sat_soil_dsl_init: parameter-file scalars) */
  int32_t use_z0m_snowmelt;             /* 1 with Meier2022 (namelist_defaults_ctsm.xml:624) */
  /* the infiltration chain: soilhydrology_inparm h2osfcflag (SoilHydrologyType.F90:352), crop_fsat_equals_zero
   * (SaturatedExcessRunoffMod.F90), parameter-file scalars fff (SaturatedExcessRunoffMod.F90:56), pc, mu (SurfaceWaterMod.F90:42-43) */
  int32_t h2osfcflag;                   /* 1 */
  int32_t crop_fsat_equals_zero;        /* 0 */
  int32_t reserved_i[2];
  double  zlnd, zsno, zglc, d_max, frac_sat_soil_dsl_init;
  double  fff, pc, mu;
  /* the snow routines of HydrologyNoDrainage (SnowHydrologyMod.F90): clm_snowhydrology_inparm :188-283 (namelist), params_inst
   * :70-91 (parameter file), scf_swenson_lawrence_2012_inparm int_snow_max, clm_varctl use_subgrid_fluxes */
  int32_t snow_overburden_compaction_method;  /* 1 = Anderson1976, 2 = Vionnet2012 (clm5 / clm6) */
  int32_t wind_dependent_snow_density;        /* 1 */
  int32_t use_subgrid_fluxes;                 /* 1 */
  int32_t snicar_use_aerosol;                 /* 1: AerosolFluxes deposits forc_aer on the top snow layer (AerosolMod.F90:754) */
  double  snow_dzmin_1, snow_dzmin_2, snow_dzmax_l_1, snow_dzmax_l_2, snow_dzmax_u_1, snow_dzmax_u_2;
  double  overburden_compress_Tfactor, int_snow_max;
  double  wimp, ssi, drift_gs, eta0_anderson, eta0_vionnet, rho_max, tau_ref, ceta, snw_rds_min, upplim_destruct_metamorph;
  double  scvng_fct_mlt_sf, scvng_fct_mlt_bcphi, scvng_fct_mlt_bcpho, scvng_fct_mlt_dst1, scvng_fct_mlt_dst2,
          scvng_fct_mlt_dst3, scvng_fct_mlt_dst4;
  /* SnowCapping: clm_varcon h2osno_max (namelist, 10000 mm for the standard structure), clm_snowhydrology_inparm reset_snow /
   * reset_snow_glc / reset_snow_glc_ela (SnowHydrologyMod.F90:160-179) */
  double  h2osno_max, reset_snow_glc_ela;
  int32_t reset_snow, reset_snow_glc;
} ctsm_params_t;

typedef struct ctsm_b200_ctx ctsm_b200_ctx;

/* ---- field structs generated from ctsm_b200_fields.def ------------------- */
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) ctype* name;
typedef struct ctsm_soilwater_fields_t {
  ctsm_bounds_t alloc;   /* lower/upper bounds the arrays were allocated with */
#define CTSM_FIELDS_SOILWATER
#include "ctsm_b200_fields.def"
#undef CTSM_FIELDS_SOILWATER
} ctsm_soilwater_fields_t;

typedef struct ctsm_soiltemperature_fields_t {
  ctsm_bounds_t alloc;
#define CTSM_FIELDS_SOILTEMPERATURE
#include "ctsm_b200_fields.def"
#undef CTSM_FIELDS_SOILTEMPERATURE
} ctsm_soiltemperature_fields_t;

typedef struct ctsm_canopyfluxes_fields_t {
  ctsm_bounds_t alloc;
#define CTSM_FIELDS_CANOPYFLUXES
#include "ctsm_b200_fields.def"
#undef CTSM_FIELDS_CANOPYFLUXES
} ctsm_canopyfluxes_fields_t;

typedef struct ctsm_plantsink_fields_t {
  ctsm_bounds_t alloc;
#define CTSM_FIELDS_PLANTSINK
#include "ctsm_b200_fields.def"
#undef CTSM_FIELDS_PLANTSINK
} ctsm_plantsink_fields_t;

typedef struct ctsm_plantsinkdefault_fields_t {
  ctsm_bounds_t alloc;
#define CTSM_FIELDS_PLANTSINKDEFAULT
#include "ctsm_b200_fields.def"
#undef CTSM_FIELDS_PLANTSINKDEFAULT
} ctsm_plantsinkdefault_fields_t;

typedef struct ctsm_preflux_fields_t {
  ctsm_bounds_t alloc;
#define CTSM_FIELDS_PREFLUX
#include "ctsm_b200_fields.def"
#undef CTSM_FIELDS_PREFLUX
} ctsm_preflux_fields_t;

typedef struct ctsm_surfacehumidity_fields_t {
  ctsm_bounds_t alloc;
#define CTSM_FIELDS_SURFACEHUMIDITY
#include "ctsm_b200_fields.def"
#undef CTSM_FIELDS_SURFACEHUMIDITY
} ctsm_surfacehumidity_fields_t;

typedef struct ctsm_baregroundfluxes_fields_t {
  ctsm_bounds_t alloc;
#define CTSM_FIELDS_BAREGROUNDFLUXES
#include "ctsm_b200_fields.def"
#undef CTSM_FIELDS_BAREGROUNDFLUXES
} ctsm_baregroundfluxes_fields_t;

typedef struct ctsm_infiltration_fields_t {
  ctsm_bounds_t alloc;
#define CTSM_FIELDS_INFILTRATION
#include "ctsm_b200_fields.def"
#undef CTSM_FIELDS_INFILTRATION
} ctsm_infiltration_fields_t;

typedef struct ctsm_ozone_fields_t {
  ctsm_bounds_t alloc;
#define CTSM_FIELDS_OZONE
#include "ctsm_b200_fields.def"
#undef CTSM_FIELDS_OZONE
} ctsm_ozone_fields_t;

typedef struct ctsm_snowwater_fields_t {
  ctsm_bounds_t alloc;
#define CTSM_FIELDS_SNOWWATER
#include "ctsm_b200_fields.def"
#undef CTSM_FIELDS_SNOWWATER
} ctsm_snowwater_fields_t;

typedef struct ctsm_snowlayers_fields_t {
  ctsm_bounds_t alloc;
#define CTSM_FIELDS_SNOWLAYERS
#include "ctsm_b200_fields.def"
#undef CTSM_FIELDS_SNOWLAYERS
} ctsm_snowlayers_fields_t;

typedef struct ctsm_snowcapping_fields_t {
  ctsm_bounds_t alloc;
#define CTSM_FIELDS_SNOWCAPPING
#include "ctsm_b200_fields.def"
#undef CTSM_FIELDS_SNOWCAPPING
} ctsm_snowcapping_fields_t;

typedef struct ctsm_watertable_fields_t {
  ctsm_bounds_t alloc;
#define CTSM_FIELDS_WATERTABLE
#include "ctsm_b200_fields.def"
#undef CTSM_FIELDS_WATERTABLE
} ctsm_watertable_fields_t;

typedef struct ctsm_hydrodiag_fields_t {
  ctsm_bounds_t alloc;
#define CTSM_FIELDS_HYDRODIAG
#include "ctsm_b200_fields.def"
#undef CTSM_FIELDS_HYDRODIAG
} ctsm_hydrodiag_fields_t;

typedef struct ctsm_soilfluxes_fields_t {
  ctsm_bounds_t alloc;
#define CTSM_FIELDS_SOILFLUXES
#include "ctsm_b200_fields.def"
#undef CTSM_FIELDS_SOILFLUXES
} ctsm_soilfluxes_fields_t;

typedef struct ctsm_patch2col_fields_t {
  ctsm_bounds_t alloc;
#define CTSM_FIELDS_PATCH2COL
#include "ctsm_b200_fields.def"
#undef CTSM_FIELDS_PATCH2COL
} ctsm_patch2col_fields_t;

typedef struct ctsm_waterbalance_fields_t {
  ctsm_bounds_t alloc;
#define CTSM_FIELDS_WATERBALANCE
#include "ctsm_b200_fields.def"
#undef CTSM_FIELDS_WATERBALANCE
} ctsm_waterbalance_fields_t;

typedef struct ctsm_watergridbalance_fields_t {
  ctsm_bounds_t alloc;
#define CTSM_FIELDS_WATERGRIDBALANCE
#include "ctsm_b200_fields.def"
#undef CTSM_FIELDS_WATERGRIDBALANCE
} ctsm_watergridbalance_fields_t;

typedef struct ctsm_balancecheck_fields_t {
  ctsm_bounds_t alloc;
#define CTSM_FIELDS_BALANCECHECK
#include "ctsm_b200_fields.def"
#undef CTSM_FIELDS_BALANCECHECK
} ctsm_balancecheck_fields_t;
#undef CTSM_F

/* result of BalanceCheck / EnergyBalanceCheck: clump-local maxval / maxloc of each residual
 * (BalanceCheckMod.F90:610,696,806,1009,1047,1066,1101) and the warning / abort decisions */
enum { CTSM_BAL_H2O_COL = 0, CTSM_BAL_H2O_GRC = 1, CTSM_BAL_H2OSNO = 2, CTSM_BAL_SOL = 3, CTSM_BAL_LON = 4,
       CTSM_BAL_SEB = 5, CTSM_BAL_SOI = 6, CTSM_BAL_NKIND = 7 };
typedef struct ctsm_balance_report_t {
  double  max_abs[CTSM_BAL_NKIND];   /* maxval(abs(err)) over the bounds */
  int32_t index[CTSM_BAL_NKIND];     /* maxloc: lowest 1-based index attaining it (0 when nothing qualifies) */
  int32_t warn[CTSM_BAL_NKIND];      /* the reference would write(iulog) a WARNING */
  int32_t abort_kind;                /* first kind (in the reference's order of checks) that would endrun, or -1 */
  int32_t skip_steps;
} ctsm_balance_report_t;

/* setFiltersOneGroup (filterMod.F90:303-592): the index lists of `type clumpfilter` (filterMod.F90:29-115) that depend
 * on subgrid type and "active" status only.  Every list is a stable, ascending-index compaction (bit-exact target).
 * Lists the reference leaves empty under the given switches (bgc_soilc, bgc_vegp, pcropp, soilnopcropp) get num = 0. */
enum { CTSM_FLT_ALLC = 0, CTSM_FLT_LAKEC, CTSM_FLT_NOLAKEC, CTSM_FLT_BGC_SOILC, CTSM_FLT_SOILC, CTSM_FLT_HYDROLOGYC,
       CTSM_FLT_URBANC, CTSM_FLT_NOURBANC, CTSM_FLT_ICEC, CTSM_FLT_DO_SMB_C,                    /* column lists */
       CTSM_FLT_LAKEP, CTSM_FLT_NOLAKEP, CTSM_FLT_NOLAKEURBANP, CTSM_FLT_BGC_VEGP, CTSM_FLT_SOILP, CTSM_FLT_PCROPP,
       CTSM_FLT_SOILNOPCROPP, CTSM_FLT_URBANP, CTSM_FLT_NOURBANP,                              /* patch lists */
       CTSM_FLT_URBANL, CTSM_FLT_NOURBANL,                                                     /* landunit lists */
       CTSM_FLT_COUNT };
typedef struct ctsm_filter_inputs_t {
  ctsm_bounds_t alloc;                       /* bounds the arrays below are allocated with (lower bounds) */
  const int32_t* col_active;                 /* col%active(begc:endc), 0/1 */
  const int32_t* col_landunit;               /* col%landunit, 1-based proc-local landunit index */
  const int32_t* col_gridcell;               /* col%gridcell */
  const int32_t* col_hydrologically_active;  /* col%hydrologically_active */
  const int32_t* lun_active;                 /* lun%active(begl:endl) */
  const int32_t* lun_lakpoi;                 /* lun%lakpoi */
  const int32_t* lun_urbpoi;                 /* lun%urbpoi */
  const int32_t* lun_itype;                  /* lun%itype */
  const int32_t* patch_active;               /* patch%active(begp:endp) */
  const int32_t* patch_landunit;             /* patch%landunit */
  const int32_t* patch_itype;                /* patch%itype */
  const int32_t* melt_replaced_by_ice_grc;   /* glc_behavior%melt_replaced_by_ice_grc(begg:endg) */
  int32_t include_inactive, use_cn, use_fates, use_fates_bgc;
  int32_t npcropmin, npcropmax;              /* is_prognostic_crop (pftconMod.F90:1812-1825) */
} ctsm_filter_inputs_t;
typedef struct ctsm_filters_t {
  int32_t* list[CTSM_FLT_COUNT];             /* each sized to the extent of its subgrid level in `bounds` */
  int32_t  num[CTSM_FLT_COUNT];              /* out: list lengths (host memory in every mode) */
} ctsm_filters_t;

/* ---- lifecycle ------------------------------------------------------------ */
void ctsm_b200_default_params(ctsm_params_t* p);
int  ctsm_b200_init(const ctsm_params_t* p, ctsm_b200_ctx** ctx);
int  ctsm_b200_finalize(ctsm_b200_ctx* ctx);
/* wait for the context stream; fills *st from the device-side first-failure record */
int  ctsm_b200_sync(ctsm_b200_ctx* ctx, ctsm_status_t* st);
/* the cudaStream_t (as void*) all kernels of this context are launched on */
void* ctsm_b200_stream(ctsm_b200_ctx* ctx);
/* the context's other streams (for timelines): kind 0 = compute (same as ctsm_b200_stream), 1 = CanopyFluxes tail kernel,
 * 2 = host-to-device copies of a resident window, 3 = device-to-host copies (2 and 3 exist after the first window) */
void* ctsm_b200_stream_of(ctsm_b200_ctx* ctx, int kind);
/* number of kernel launches issued by this context so far */
int64_t ctsm_b200_launch_count(const ctsm_b200_ctx* ctx);
/* "cudaErrorName at file:line: text" of the last CUDA runtime error this process's library calls met ("" if none) */
const char* ctsm_b200_last_cuda_error(void);
/* Perturbed-parameter ensembles (BASELINE.json config 5: members batched as independent columns).  The PFT tables carry the
 * members through ctsm_params_t.npft_table (patch%itype = member*(mxpft+1) + pft); this call does the same for the SCALAR
 * parameters the survey lists: per-member values (host arrays of length nmember, NULL = keep the scalar of ctsm_params_t) of
 * e_ice (SoilWaterMovementMod.F90:34), csoilc, cv, a_coef, z_dl (CanopyFluxesMod.F90:61-70).  A patch's member is
 * itype / (mxpft+1); col_member(begc:endc) gives every column's member (0-based; needed for e_ice).  nmember must equal
 * npft_table / (mxpft+1); nmember = 0 clears.  Synchronous, the arrays are copied. */
int  ctsm_b200_set_member_params(ctsm_b200_ctx* ctx, int nmember, const double* e_ice, const double* csoilc, const double* cv,
                                 const double* a_coef, const double* z_dl, const int32_t* col_member, int begc, int endc);
/* Scheduling knobs of CanopyFluxes' ITERATION loop (results do not depend on them; defaults come from the environment
 * variables CTSM_B200_TAIL_MAX / CTSM_B200_NT_BUDGET / CTSM_B200_TAIL_LANES, else 0 / 0 / 1 = tail kernel off: on B200 the
 * per-patch nested-loop kernel is instruction-fetch bound and only pays for calls of a few patches):
 *   tail_max    once a pass has at most this many unconverged patches they all leave the list-driven bulk rounds and
 *               one thread per patch runs each through its remaining passes (0: never);
 *   nt_budget   a calcstress solve (PhotosynthesisMod.F90:4579) still running after this many Newton iterations sends
 *               its patch to the same tail kernel (0: never);
 *   tail_lanes  patches carried by one warp of the tail kernel (1..32);
 *   nt_split    1 (default; CTSM_B200_NT_SPLIT): large calcstress queues run prologue / Newton iterations / epilogue as three
 *               kernels, 0: as lane tasks of one persistent kernel.
 * A negative argument keeps the current value. */
int  ctsm_b200_set_tuning(ctsm_b200_ctx* ctx, int tail_max, int nt_budget, int tail_lanes, int nt_split);
/* SoilTemperature formulation (results are bit-identical): soil_stream 1 (default; CTSM_B200_SOIL_STREAM) streams the levels
 * through registers and a coalesced scratch (soiltemp_stream_kernel), 0 keeps the per-thread level arrays (soiltemp_kernel);
 * a negative value leaves the setting unchanged. */
int  ctsm_b200_set_soil_tuning(ctsm_b200_ctx* ctx, int soil_stream);
/* SoilWater's second pass (the columns whose full time step was rejected): sw_warp 1 (default; CTSM_B200_SW_WARP) runs one warp
 * per column with the soil levels across the lanes (soilwater_retry_warp_kernel), 0 one thread per column (soilwater_kernel<2>).
 * Bit-identical results; a negative value leaves the setting unchanged. */
int  ctsm_b200_set_soilwater_tuning(ctsm_b200_ctx* ctx, int sw_warp);
/* Compute_EffecRootFrac_And_VertTranSink_HydStress: sink_warp 1 (default; CTSM_B200_SINK_WARP) runs one warp per column with the
 * column's patches across the lanes (plantsink_warp_kernel: contiguous reads of k_soil_root), 0 one thread per column
 * (plantsink_kernel).  Bit-identical results; a negative value leaves the setting unchanged. */
int  ctsm_b200_set_sink_tuning(ctsm_b200_ctx* ctx, int sink_warp);
/* Diagnostic of the last ctsm_b200_canopyfluxes call (synchronises the stream): for every ITERATION round r < cap,
 * list_len[r] = patches the round's list kernels served, tail_end[r] = patches handed to the tail kernel up to and
 * including round r.  Returns the number of rounds written. */
int  ctsm_b200_canopy_round_stats(ctsm_b200_ctx* ctx, int32_t* list_len, int32_t* tail_end, int cap);
/* Resident window for CTSM_MEM_HOST callers (SURVEY.md 8b "Ownership": the library owns device mirrors keyed by host
 * base address).  Between _begin and _end the host must not touch the arrays it passes to hot-path calls.  In the window
 * every host array has a persistent device mirror (filled once, at its first use ever); a call uploads only the parts
 * of its IN / INOUT fields that no earlier call of the window uploaded or produced, never an OUT field; calls return
 * once their work is queued (status comes from _end); each call's OUT / INOUT fields are downloaded over its bounds
 * as soon as its kernels finish.  Copies run on their own streams: when the host issues the step clump after clump
 * (as clm_drv's clump loop does), uploads, kernels and downloads of successive clumps overlap.
 * Contract: an array the window only writes (OUT) keeps, outside the filters, the values it had when its mirror was
 * created or last downloaded; host code that changes such an array between windows calls ctsm_b200_host_invalidate.
 * ctsm_b200_host_window_bytes reports the bytes the last (or current) window moved in each direction. */
int  ctsm_b200_host_window_begin(ctsm_b200_ctx* ctx);
int  ctsm_b200_host_window_end(ctsm_b200_ctx* ctx, ctsm_status_t* st);
int  ctsm_b200_host_window_bytes(const ctsm_b200_ctx* ctx, uint64_t* h2d_bytes, uint64_t* d2h_bytes);
/* drop the device mirror of one host array (NULL: of all); not inside a window */
int  ctsm_b200_host_invalidate(ctsm_b200_ctx* ctx, const void* host_ptr);
/* page-lock a host array so that CTSM_MEM_HOST staging runs at full PCIe rate */
int  ctsm_b200_host_register(void* ptr, uint64_t bytes);
int  ctsm_b200_host_unregister(void* ptr);
const char* ctsm_b200_version(void);

/* ---- numerical primitives -------------------------------------------------- */

/* Tridiagonal(bounds, lbj, ubj, jtop, numf, filter, a, b, c, r, u): TridiagonalMod.F90:23-91.
 * a,b,c,r,u are (begc:endc, lbj:ubj); jtop is (begc:endc); filter(1:numf) holds
 * column indices in [begc,endc].  Non-pivoting Thomas algorithm, only levels
 * j >= jtop(ci) participate. */
int ctsm_b200_tridiagonal(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int lbj, int ubj,
                          const int32_t* jtop, int numf, const int32_t* filter,
                          const double* a, const double* b, const double* c, const double* r,
                          double* u, int mem);

/* BandDiagonal(bounds, lbj, ubj, jtop, jbot, numf, filter, nband, b, r, u):
 * BandDiagonalMod.F90:29-221.  b is (begc:endc, nband, lbj:ubj) with nband = 5;
 * per column the rows jtop(ci)..jbot(ci) are solved with the semantics of
 * LAPACK dgbsv(n, kl=2, ku=2, nrhs=1) (partial pivoting, dgbtf2 + dgbtrs).
 * info != 0 is reported like the reference's endrun (CTSM_ERR_DGBSV). */
int ctsm_b200_banddiagonal(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int lbj, int ubj,
                           const int32_t* jtop, const int32_t* jbot, int numf, const int32_t* filter,
                           int nband, const double* b, const double* r, double* u,
                           int mem, ctsm_status_t* st);

/* Batched LAPACK dgtsv(n, nrhs=1) as called at SoilWaterMovementMod.F90:1279-1299:
 * for every filter column ci, n = nlayers(ci); dl = amx(ci,2:n), d = bmx(ci,1:n),
 * du = cmx(ci,1:n-1), rhs = rmx(ci,1:n); x(ci,1:n) receives the solution.
 * All 2-D arrays are (begc:endc, 1:nlev). */
int ctsm_b200_dgtsv_batch(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int nlev,
                          const int32_t* nlayers, int numf, const int32_t* filter,
                          const double* amx, const double* bmx, const double* cmx, const double* rmx,
                          double* x, int mem, ctsm_status_t* st);

/* ---- physics ---------------------------------------------------------------- */

/* SoilWater(bounds, num_hydrologyc, filter_hydrologyc, num_urbanc, filter_urbanc, ...):
 * SoilWaterMovementMod.F90:240-243 with soilwater_movement_method = moisture_form (:976). */
int ctsm_b200_soilwater(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds,
                        int num_hydrologyc, const int32_t* filter_hydrologyc,
                        const ctsm_soilwater_fields_t* f, int mem, ctsm_status_t* st);

/* SoilTemperature(bounds, num_urbanl, filter_urbanl, num_urbanc, filter_urbanc,
 *                 num_nolakep, filter_nolakep, num_nolakec, filter_nolakec, ...):
 * SoilTemperatureMod.F90:92-95.  Urban filters must be empty (CTSM_ERR_URBAN). */
int ctsm_b200_soiltemperature(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds,
                              int num_nolakep, const int32_t* filter_nolakep,
                              int num_nolakec, const int32_t* filter_nolakec,
                              const ctsm_soiltemperature_fields_t* f, int mem, ctsm_status_t* st);

/* CanopyFluxes(bounds, num_exposedvegp, filter_exposedvegp, ...): CanopyFluxesMod.F90:191-198
 * with use_hydrstress=.true. (PhotosynthesisHydraulicStress, PhotosynthesisMod.F90:2704).
 * Failures reported like the reference's endrun sites: CTSM_ERR_FORC_HGT (:997-1002),
 * CTSM_ERR_GS_NEG, CTSM_ERR_BRENT, CTSM_ERR_QUADRATIC.  st->n_warnings counts patches
 * whose canopy energy balance error exceeds 0.1 W/m2 (:1746-1760) plus ustar*thvstar>0
 * occurrences (:1408-1423). */
int ctsm_b200_canopyfluxes(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds,
                           int num_exposedvegp, const int32_t* filter_exposedvegp,
                           const ctsm_canopyfluxes_fields_t* f, int mem, ctsm_status_t* st);

/* setFiltersOneGroup(bounds, this_filter, include_inactive, glc_behavior): filterMod.F90:303.  Synchronous (the counts
 * are returned to the host).  mem says where the input arrays and the output lists live. */
int ctsm_b200_set_filters(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, const ctsm_filter_inputs_t* in,
                          ctsm_filters_t* out, int mem);

/* setExposedvegpFilter(bounds, frac_veg_nosno): filterMod.F90:595-648.  Order-preserving
 * split of filter_nolakeurbanp into exposedvegp (frac_veg_nosno > 0) / noexposedvegp.
 * Output lists must hold num_nolakeurbanp entries; counts are returned through the
 * pointers (host memory in every mode; DEVICE mode synchronises to return them). */
int ctsm_b200_set_exposedvegp_filter(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds,
                                     int num_nolakeurbanp, const int32_t* filter_nolakeurbanp,
                                     const int32_t* frac_veg_nosno /* (begp:endp) */,
                                     int32_t* filter_exposedvegp, int32_t* num_exposedvegp,
                                     int32_t* filter_noexposedvegp, int32_t* num_noexposedvegp, int mem);

/* Compute_EffecRootFrac_And_VertTranSink_HydStress(bounds, num_filterc, filterc, ...):
 * SoilWaterPlantSinkMod.F90:236-328 (the PHS sink term SoilWater consumes). */
int ctsm_b200_vert_tran_sink_hydstress(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds,
                                       int num_filterc, const int32_t* filterc,
                                       const ctsm_plantsink_fields_t* f, int mem, ctsm_status_t* st);

/* BiogeophysPreFluxCalcs(bounds, num_nolakec, filter_nolakec, num_nolakep, filter_nolakep, num_urbanc, filter_urbanc, ...):
 * BiogeophysPreFluxCalcsMod.F90:58-118, call site clm_driver.F90:680.  Runs SetZ0mDisp, SetRoughnessLengthsAndForcHeightsNonLake,
 * CalcInitialTemperatureAndEnergyVars and calc_soilevap_resis.  num_urbanc must be 0 and no column of an urban landunit may
 * be in filter_nolakec (CTSM_ERR_URBAN).  time_flags carries the clock tests of SetZ0mDisp (:174-184):
 * bit 0 = is_first_step() .or. get_nstep() <= GetBalanceCheckSkipSteps()-1, bit 1 = is_beg_curr_year(). */
#define CTSM_TIME_FIRST_STEPS 1
#define CTSM_TIME_BEG_CURR_YEAR 2
int ctsm_b200_biogeophys_pre_flux_calcs(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds,
                                        int num_nolakec, const int32_t* filter_nolakec,
                                        int num_nolakep, const int32_t* filter_nolakep,
                                        int num_urbanc, const int32_t* filter_urbanc, int time_flags,
                                        const ctsm_preflux_fields_t* f, int mem, ctsm_status_t* st);

/* CalculateSurfaceHumidity(bounds, num_nolakec, filter_nolakec, ...): SurfaceHumidityMod.F90:41-239, call site
 * clm_driver.F90:702.  Urban columns are refused (CTSM_ERR_URBAN). */
int ctsm_b200_calculate_surface_humidity(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds,
                                         int num_nolakec, const int32_t* filter_nolakec,
                                         const ctsm_surfacehumidity_fields_t* f, int mem, ctsm_status_t* st);

/* BareGroundFluxes(bounds, num_noexposedvegp, filter_noexposedvegp, ...): BareGroundFluxesMod.F90:63-529, call site
 * clm_driver.F90:711.  use_lch4 = .false.; the human-stress indices follow ctsm_params_t.calc_human_stress_indices. */
int ctsm_b200_bare_ground_fluxes(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds,
                                 int num_noexposedvegp, const int32_t* filter_noexposedvegp,
                                 const ctsm_baregroundfluxes_fields_t* f, int mem, ctsm_status_t* st);

/* The call sequence HydrologyNoDrainageMod.F90:297-337 as one entry: SetSoilWaterFractions, SetFloodc (over filter_nolakec),
 * SaturatedExcessRunoff, SetQflxInputs, InfiltrationExcessRunoff, RouteInfiltrationExcess, UpdateH2osfc, Infiltration,
 * TotalSurfaceRunoff (over filter_hydrologyc).  fsat_method = TOPModel, qinmax_method = hksat (the clm5 / clm6 defaults);
 * num_urbanc must be 0 and urban columns in the filters are refused (CTSM_ERR_URBAN). */
int ctsm_b200_hydrology_infiltration(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds,
                                     int num_nolakec, const int32_t* filter_nolakec,
                                     int num_hydrologyc, const int32_t* filter_hydrologyc,
                                     int num_urbanc, const int32_t* filter_urbanc,
                                     const ctsm_infiltration_fields_t* f, int mem, ctsm_status_t* st);

/* ozone_inst%CalcOzoneUptake(bounds, num_exposedvegp, filter_exposedvegp, forc_pbot, forc_th, rssun, rssha, rb, ram, tlai, forc_o3):
 * OzoneMod.F90:356-511, the call CanopyFluxes makes at CanopyFluxesMod.F90:1690 (the host's CanopyFluxes body issues it right after
 * ctsm_b200_canopyfluxes, whose outputs rssun / rssha / rb1 / ram1 it reads).  dtime is the integer get_step_size() of the reference. */
int ctsm_b200_calc_ozone_uptake(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_exposedvegp, const int32_t* filter_exposedvegp,
                                const ctsm_ozone_fields_t* f, int mem, ctsm_status_t* st);
/* ozone_inst%CalcOzoneStress(bounds, num_exposedvegp, filter_exposedvegp, num_noexposedvegp, filter_noexposedvegp): OzoneMod.F90:514-782
 * (clm_driver.F90:690).  stress_method: 1 = Lombardozzi2015 (o3coefv / o3coefg), 2 = Falk (o3coefjmax; only when is_time_to_run_luna). */
int ctsm_b200_calc_ozone_stress(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_exposedvegp, const int32_t* filter_exposedvegp,
                                int num_noexposedvegp, const int32_t* filter_noexposedvegp, int stress_method, int is_time_to_run_luna,
                                const ctsm_ozone_fields_t* f, int mem, ctsm_status_t* st);

/* BuildSnowFilter(bounds, num_nolakec, filter_nolakec, num_snowc, filter_snowc, num_nosnowc, filter_nosnowc):
 * SnowHydrologyMod.F90:3975-4010, called at HydrologyNoDrainageMod.F90:279 and :402.  Order-preserving split of filter_nolakec
 * by col%snl < 0.  snl is the (alloc_begc : ...) array; output lists must hold num_nolakec entries; counts come back through
 * host pointers in every mode (DEVICE mode synchronises to return them). */
int ctsm_b200_build_snow_filter(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_nolakec, const int32_t* filter_nolakec,
                                const int32_t* snl, int alloc_begc, int alloc_endc,
                                int32_t* filter_snowc, int32_t* num_snowc, int32_t* filter_nosnowc, int32_t* num_nosnowc, int mem);

/* SnowWater(bounds, num_snowc, filter_snowc, num_nosnowc, filter_nosnowc, atm2lnd_inst, aerosol_inst, water_inst):
 * SnowHydrologyMod.F90:1015-1165.  Fails with CTSM_ERR_SNOW_NEGATIVE where the reference calls endrun
 * ("h2osoi_ice / h2osoi_liq has gone significantly negative", :1244-1287). */
int ctsm_b200_snow_water(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_snowc, const int32_t* filter_snowc,
                         int num_nosnowc, const int32_t* filter_nosnowc, const ctsm_snowwater_fields_t* f, int mem,
                         ctsm_status_t* st);

/* SnowCapping(bounds, num_initc, filter_initc, num_snowc, filter_snowc, topo_inst, aerosol_inst, water_inst):
 * SnowHydrologyMod.F90:3121-3247 (InitFlux_SnowCapping :3250, BulkFlux_SnowCappingFluxes :3288 with SnowCappingExcess :3397 and
 * CalculateTotalH2osno WaterStateType.F90:887-897, UpdateState_RemoveSnowCappingFluxes :3575, SnowCappingUpdateDzAndAerosols
 * :3623); HydrologyNoDrainageMod.F90:377.  nstep = get_nstep() (the snow reset of reset_snow / reset_snow_glc is active during the
 * first 4 * nlevsno steps).  Fails with CTSM_ERR_SNOW_NEGATIVE (info = 3) on "capping procedure failed (negative mass remaining)". */
int ctsm_b200_snow_capping(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_initc, const int32_t* filter_initc,
                           int num_snowc, const int32_t* filter_snowc, const ctsm_snowcapping_fields_t* f, int nstep, int mem,
                           ctsm_status_t* st);

/* SnowCompaction, CombineSnowLayers, DivideSnowLayers(is_lake = .false.), ZeroEmptySnowLayers over filter_snowc: the call
 * sequence HydrologyNoDrainageMod.F90:381-399 (SnowHydrologyMod.F90:1870, :2083, :2510, :2898).  Lake and urban columns in the
 * filter are refused (CTSM_ERR_URBAN / CTSM_ERR_BAD_ARG). */
int ctsm_b200_snow_layers(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_snowc, const int32_t* filter_snowc,
                          const ctsm_snowlayers_fields_t* f, int mem, ctsm_status_t* st);

/* PerchedWaterTable, ThetaBasedWaterTable, RenewCondensation over filter_hydrologyc: the call sequence
 * HydrologyNoDrainageMod.F90:359-373 (use_aquifer_layer = .false.; SoilHydrologyMod.F90:1525, :1933, :2569).  num_urbanc must be 0.
 * Fails with CTSM_ERR_SNOW_NEGATIVE (info = 2) where RenewCondensation calls endrun ("h2osoi_ice has gone significantly negative"). */
int ctsm_b200_water_table(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_hydrologyc, const int32_t* filter_hydrologyc,
                          int num_urbanc, const int32_t* filter_urbanc, const ctsm_watertable_fields_t* f, int mem, ctsm_status_t* st);

/* The column diagnostics HydrologyNoDrainage computes inline after its second BuildSnowFilter (HydrologyNoDrainageMod.F90:420-757).
 * filter_snowc / filter_nosnowc are the new snow filters; num_urbanc must be 0. */
int ctsm_b200_hydrology_diagnostics(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_nolakec, const int32_t* filter_nolakec,
                                    int num_snowc, const int32_t* filter_snowc, int num_nosnowc, const int32_t* filter_nosnowc,
                                    int num_hydrologyc, const int32_t* filter_hydrologyc, int num_urbanc, const int32_t* filter_urbanc,
                                    const ctsm_hydrodiag_fields_t* f, int mem, ctsm_status_t* st);

/* Compute_EffecRootFrac_And_VertTranSink_Default(bounds, num_filterc, filterc, ...): SoilWaterPlantSinkMod.F90:332-424,
 * what Compute_EffecRootFrac_And_VertTranSink (:18-142) calls for every column class when use_hydrstress = .false. */
int ctsm_b200_vert_tran_sink_default(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds,
                                     int num_filterc, const int32_t* filterc,
                                     const ctsm_plantsinkdefault_fields_t* f, int mem, ctsm_status_t* st);

/* SoilFluxes(bounds, num_urbanl, filter_urbanl, num_urbanp, filter_urbanp, num_nolakec, filter_nolakec, num_nolakep,
 * filter_nolakep, ...): SoilFluxesMod.F90:37-521, call site clm_driver.F90:921.  The urban filters are not part of this
 * hot path: a column of an urban landunit in filter_nolakec is refused with CTSM_ERR_URBAN. */
int ctsm_b200_soilfluxes(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_nolakec, const int32_t* filter_nolakec,
                         int num_nolakep, const int32_t* filter_nolakep, const ctsm_soilfluxes_fields_t* f, int mem,
                         ctsm_status_t* st);

/* clm_drv_patch2col(bounds, num_allc, filter_allc, num_nolakec, filter_nolakec, energyflux_inst, waterfluxbulk_inst):
 * src/main/clm_driver.F90:1655-1739, call site :936. */
int ctsm_b200_patch2col(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_allc, const int32_t* filter_allc,
                        int num_nolakec, const int32_t* filter_nolakec, const ctsm_patch2col_fields_t* f, int mem,
                        ctsm_status_t* st);

/* BeginWaterColumnBalance(bounds, num_nolakec, filter_nolakec, num_lakec, filter_lakec, ...): BalanceCheckMod.F90:171,
 * call site clm_driver.F90:414.  Bulk water, use_aquifer_layer = .false. (clm5/clm6 default).  Non-lake columns get
 * ComputeWaterMassNonLake, lake columns ComputeWaterMassLake without the lake water itself (TotalWaterAndHeatMod.F90:144,
 * 395-460: snow and soil layers only), both get h2osno_old.  aquifer_water_baseline: waterstate_inst scalar. */
int ctsm_b200_begin_water_column_balance(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_nolakec,
                                         const int32_t* filter_nolakec, int num_lakec, const int32_t* filter_lakec,
                                         const ctsm_waterbalance_fields_t* f, double aquifer_water_baseline, int mem,
                                         ctsm_status_t* st);

/* WaterGridcellBalance(bounds, num_nolakec, filter_nolakec, num_lakec, filter_lakec, water_inst, lakestate_inst,
 * use_aquifer_layer, flag): BalanceCheckMod.F90:132-350, call sites clm_driver.F90:331 ('begwb') and :1411 ('endwb').
 * Bulk water, use_aquifer_layer = .false., hillslope routing off.  flag_endwb = 0 writes begwb_grc, 1 writes endwb_grc.
 * The two dribbler amounts are inputs (the dribblers themselves belong to dynamic land cover, outside the hot path).
 * c2g failure (sum of column weights > 1) is reported like the reference's endrun: CTSM_ERR_BALANCE at the gridcell. */
int ctsm_b200_water_gridcell_balance(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_nolakec,
                                     const int32_t* filter_nolakec, int num_lakec, const int32_t* filter_lakec,
                                     const ctsm_watergridbalance_fields_t* f, double aquifer_water_baseline, int flag_endwb,
                                     int mem, ctsm_status_t* st);

/* BalanceCheckInit(): BalanceCheckMod.F90:74-95; skip_steps = max(2, nint(3600/dtime)) + 1.  Returns skip_steps. */
int ctsm_b200_balancecheck_init(ctsm_b200_ctx* ctx);

/* BalanceCheck(bounds, num_allc, filter_allc, ...) including EnergyBalanceCheck:
 * BalanceCheckMod.F90:445-857, 859-1119.  DAnstep = get_nstep_since_startup_or_lastDA_restart_or_pause().
 * CTSM_MEM_HOST outside a window: synchronous; returns CTSM_ERR_BALANCE and fills st like endrun(subgrid_index=,
 * subgrid_level=) when a residual exceeds its abort threshold after the skip steps.  CTSM_MEM_DEVICE, and host arrays
 * inside a resident window: asynchronous; *report (which must stay valid until then) is filled and the same decision
 * is returned by the next ctsm_b200_sync / ctsm_b200_host_window_end.  CTSM_ERR_BAD_ARG when called before
 * ctsm_b200_balancecheck_init (the reference aborts likewise). */
int ctsm_b200_balancecheck(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_allc, const int32_t* filter_allc,
                           const ctsm_balancecheck_fields_t* f, int DAnstep, int mem,
                           ctsm_balance_report_t* report, ctsm_status_t* st);

/* Device address of the 7 maxima (doubles, |residual|, order CTSM_BAL_*) the last asynchronous BalanceCheck call of this
 * context left, valid in stream order on ctsm_b200_stream: lets a multi-GPU host reduce the global figures with NCCL
 * (MAX over ranks) without a host round trip.  NULL before the first such call. */
void* ctsm_b200_balance_device_maxima(ctsm_b200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* CTSM_B200_H */
