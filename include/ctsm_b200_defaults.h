/* ctsm_b200_defaults.h — the clm6_0 default values of ctsm_params_t, in ONE place.
 *
 * Included by the product library (ctsm_b200_default_params, ctsm_b200/csrc/abi.cu) and by the CPU oracle
 * (oracle_default_params), so that bench.py's reference arm gets its parameters without loading the CUDA library.
 * tests/test_namelist_defaults.py pins the values against bld/namelist_files/namelist_defaults_ctsm.xml. */
#ifndef CTSM_B200_DEFAULTS_H
#define CTSM_B200_DEFAULTS_H
#include <string.h>
#include "ctsm_b200.h"

static inline void ctsm_default_params_fill(ctsm_params_t* p) {
  // clm6_0 defaults: bld/namelist_files/namelist_defaults_ctsm.xml:471-488 (soilwater_movement),
  // :511 (nlevsno), :254 (20SL_8.5m), :556 (Sturm1997); e_ice is a parameter-file scalar.
  memset(p, 0, sizeof *p);
  p->abi_version = CTSM_B200_ABI_VERSION;
  p->device = 0;
  p->nlevsno = CTSM_NLEVSNO; p->nlevgrnd = CTSM_NLEVGRND; p->nlevsoi = CTSM_NLEVSOI;
  p->dtime = 1800.0;
  p->upper_boundary_condition = 1;
  p->lower_boundary_condition = 2;
  p->flux_calculation = 1;
  p->dtmin = 60.0; p->verySmall = 1.e-8; p->xTolerUpper = 1.e-1; p->xTolerLower = 1.e-2;
  p->e_ice = 6.0;
  p->snow_thermal_cond_method = 2;
  p->snow_thermal_cond_glc_method = 2;   // clm6_0: Sturm1997 as well (namelist_defaults_ctsm.xml:559)
  // canopyfluxes_inparm / clm6_0 switches: namelist_defaults_ctsm.xml:462,454,457,622,270,726,635,108,95,104,2182
  p->itmax_canopy_fluxes = 40;
  p->use_undercanopy_stability = 0;
  p->use_biomass_heat_storage = 1;
  p->z0param_method = 2;
  p->soil_resis_method = 1;
  p->use_hydrstress = 1;
  p->use_luna = 1;
  p->stomatalcond_mtd = 2;
  p->light_inhibit = 1;
  p->modifyphoto_and_lmr_forcrop = 1;
  // parameter-file scalars: values documented inside the reference where available, otherwise
  // synthetic choices (SURVEY.md Appendix D marks which)
  p->lai_dl = 0.5; p->z_dl = 0.05; p->a_coef = 0.13; p->a_exp = 0.45; p->csoilc = 0.004; p->cv = 0.01;
  p->wind_min = 1.0;
  p->zetamaxstable = 2.0;
  p->leaf_mr_vcm = 0.015;
  p->act25 = 60.0; p->fnr = 7.16; p->cp25_yr2000 = 42.75e-6; p->kc25_coef = 404.9e-6; p->ko25_coef = 278.4e-3;
  p->fnps = 0.15; p->theta_psii = 0.7; p->theta_ip = 0.95;
  p->vcmaxha = 72000.0; p->jmaxha = 50000.0; p->tpuha = 72000.0; p->lmrha = 46390.0;
  p->kcha = 79430.0; p->koha = 36380.0; p->cpha = 37830.0;
  p->vcmaxhd = 200000.0; p->jmaxhd = 200000.0; p->tpuhd = 200000.0; p->lmrhd = 150650.0; p->lmrse = 490.0;
  p->tpu25ratio = 0.167; p->kp25ratio = 20000.0;
  p->vcmaxse_sf = 1.0; p->jmaxse_sf = 1.0; p->tpuse_sf = 1.0; p->jmax25top_sf = 1.0;
  p->calc_human_stress_indices = 1;      // FAST: namelist_defaults_ctsm.xml:235
  // roughness lengths and the dry-surface-layer parameters are parameter-file scalars (the file is not in the source tree):
  // the ctsm5.1+ parameter-file values are used (synthetic choice)
  p->use_z0m_snowmelt = 1;                 // namelist_defaults_ctsm.xml:624 (z0param_method = Meier2022)
  p->zlnd = 0.000775; p->zsno = 0.00085; p->zglc = 0.00230000005;
  p->d_max = 15.0; p->frac_sat_soil_dsl_init = 0.8;
  p->h2osfcflag = 1;                       // SoilHydrologyType.F90:352-360
  p->crop_fsat_equals_zero = 0;            // namelist_defaults_ctsm.xml (saturated_excess_runoff_inparm)
  p->fff = 0.5; p->pc = 0.4; p->mu = 0.13889;   // parameter-file scalars (clm5 technical note values; synthetic choice)
  // snow: namelist_defaults_ctsm.xml:521-544, 694 (clm6_0); parameter-file scalars: the ctsm5.1+ parameter-file values
  // (synthetic choice; scavenging factors as in Flanner et al. 2007 / SnowHydrologyMod.F90:131-132)
  p->snow_overburden_compaction_method = 2; p->wind_dependent_snow_density = 1; p->use_subgrid_fluxes = 1; p->snicar_use_aerosol = 1;
  p->snow_dzmin_1 = 0.010; p->snow_dzmin_2 = 0.015; p->snow_dzmax_l_1 = 0.03; p->snow_dzmax_l_2 = 0.07;
  p->snow_dzmax_u_1 = 0.02; p->snow_dzmax_u_2 = 0.05;
  p->overburden_compress_Tfactor = 0.08; p->int_snow_max = 2000.0;
  p->wimp = 0.05; p->ssi = 0.033; p->drift_gs = 0.35e-3; p->eta0_anderson = 9.0e5; p->eta0_vionnet = 7.62237e6;
  p->rho_max = 350.0; p->tau_ref = 172800.0; p->ceta = 250.0; p->snw_rds_min = 54.526; p->upplim_destruct_metamorph = 175.0;
  p->scvng_fct_mlt_sf = 1.0; p->scvng_fct_mlt_bcphi = 0.20; p->scvng_fct_mlt_bcpho = 0.03;
  p->scvng_fct_mlt_dst1 = 0.02; p->scvng_fct_mlt_dst2 = 0.02; p->scvng_fct_mlt_dst3 = 0.01; p->scvng_fct_mlt_dst4 = 0.01;
  p->h2osno_max = 10000.0; p->reset_snow = 0; p->reset_snow_glc = 0; p->reset_snow_glc_ela = 1.e9;   // namelist_defaults_ctsm.xml:517,546-551
  p->balance_skip_steps = -1;
  p->npft_table = CTSM_MXPFT + 1;
}

#endif /* CTSM_B200_DEFAULTS_H */
