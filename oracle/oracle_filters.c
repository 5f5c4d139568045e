/* oracle_filters.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of setFiltersOneGroup, src/main/filterMod.F90:303-592: one sequential loop per list, as the reference
 * writes them.  Semantics pinned by the reference's test_filter_col.pf (stable ascending order of a logical-array
 * filter) through tests/test_oracle_golden.py; the lists themselves have one right answer (integer work, bit-exact).
 */
#include <string.h>
#include "oracle.h"

int oracle_set_filters(const ctsm_bounds_t* b, const ctsm_filter_inputs_t* in, ctsm_filters_t* out) {
  const int begc0 = in->alloc.begc, begl0 = in->alloc.begl, begp0 = in->alloc.begp, begg0 = in->alloc.begg;
  const int inc = in->include_inactive;
#define CA(c) in->col_active[(c) - begc0]
#define CL(c) in->col_landunit[(c) - begc0]
#define PA(p) in->patch_active[(p) - begp0]
#define PL(p) in->patch_landunit[(p) - begp0]
#define LAK(l) in->lun_lakpoi[(l) - begl0]
#define URB(l) in->lun_urbpoi[(l) - begl0]
#define LT(l) in->lun_itype[(l) - begl0]
#define ADD(k, v) out->list[k][out->num[k]++] = (v)
  memset(out->num, 0, sizeof out->num);
  for (int c = b->begc; c <= b->endc; ++c) if (CA(c) || inc) ADD(CTSM_FLT_ALLC, c);                     /* :342-349 */
  for (int c = b->begc; c <= b->endc; ++c)                                                              /* :353-367 */
    if (CA(c) || inc) { if (LAK(CL(c))) ADD(CTSM_FLT_LAKEC, c); else ADD(CTSM_FLT_NOLAKEC, c); }
  for (int p = b->begp; p <= b->endp; ++p)                                                              /* :371-391 */
    if (PA(p) || inc) {
      const int l = PL(p);
      if (LAK(l)) ADD(CTSM_FLT_LAKEP, p);
      else { ADD(CTSM_FLT_NOLAKEP, p); if (!URB(l)) ADD(CTSM_FLT_NOLAKEURBANP, p); }
    }
  if (in->use_cn || in->use_fates_bgc)                                                                  /* :395-407 */
    for (int c = b->begc; c <= b->endc; ++c)
      if (CA(c) || inc) { const int t = LT(CL(c)); if (t == CTSM_ISTSOIL || t == CTSM_ISTCROP) ADD(CTSM_FLT_BGC_SOILC, c); }
  if (in->use_cn)                                                                                       /* :411-426 */
    for (int p = b->begp; p <= b->endp; ++p)
      if (PA(p) || inc) { const int t = LT(PL(p)); if (t == CTSM_ISTSOIL || t == CTSM_ISTCROP) ADD(CTSM_FLT_BGC_VEGP, p); }
  for (int c = b->begc; c <= b->endc; ++c)                                                              /* :431-441 */
    if (CA(c) || inc) { const int t = LT(CL(c)); if (t == CTSM_ISTSOIL || t == CTSM_ISTCROP) ADD(CTSM_FLT_SOILC, c); }
  for (int p = b->begp; p <= b->endp; ++p)                                                              /* :448-459 */
    if (PA(p) || inc) { const int t = LT(PL(p)); if (t == CTSM_ISTSOIL || t == CTSM_ISTCROP) ADD(CTSM_FLT_SOILP, p); }
  for (int c = b->begc; c <= b->endc; ++c)                                                              /* :463-472 */
    if (CA(c) || inc) if (in->col_hydrologically_active[c - begc0]) ADD(CTSM_FLT_HYDROLOGYC, c);
  for (int p = b->begp; p <= b->endp; ++p)                                                              /* :477-496 */
    if (!in->use_fates)
      if (PA(p) || inc) {
        const int ivt = in->patch_itype[p - begp0];
        if (ivt >= in->npcropmin && ivt <= in->npcropmax) ADD(CTSM_FLT_PCROPP, p);
        else { const int t = LT(PL(p)); if (t == CTSM_ISTSOIL || t == CTSM_ISTCROP) ADD(CTSM_FLT_SOILNOPCROPP, p); }
      }
  for (int l = b->begl; l <= b->endl; ++l)                                                              /* :500-514 */
    if (in->lun_active[l - begl0] || inc) { if (URB(l)) ADD(CTSM_FLT_URBANL, l); else ADD(CTSM_FLT_NOURBANL, l); }
  for (int c = b->begc; c <= b->endc; ++c)                                                              /* :518-533 */
    if (CA(c) || inc) { if (URB(CL(c))) ADD(CTSM_FLT_URBANC, c); else ADD(CTSM_FLT_NOURBANC, c); }
  for (int p = b->begp; p <= b->endp; ++p)                                                              /* :537-552 */
    if (PA(p) || inc) { if (URB(PL(p))) ADD(CTSM_FLT_URBANP, p); else ADD(CTSM_FLT_NOURBANP, p); }
  for (int c = b->begc; c <= b->endc; ++c)                                                              /* :554-563 */
    if (CA(c) || inc) if (LT(CL(c)) == CTSM_ISTICE) ADD(CTSM_FLT_ICEC, c);
  for (int c = b->begc; c <= b->endc; ++c)                                                              /* :565-585 */
    if (CA(c) || inc) {
      const int t = LT(CL(c));
      if (in->melt_replaced_by_ice_grc[in->col_gridcell[c - begc0] - begg0] && (t == CTSM_ISTICE || t == CTSM_ISTSOIL))
        ADD(CTSM_FLT_DO_SMB_C, c);
    }
  return 0;
}
