"""Random subgrid topology for the setFiltersOneGroup tests (filterMod.F90:303-592) and a numpy statement of the lists."""
import ctypes as C

import numpy as np

from ctsm_b200 import abi

ISTSOIL, ISTCROP, ISTICE, ISTDLAK, ISTWET, URB0 = 1, 2, 4, 5, 6, 7


def random_topology(ng, seed, p_inactive=0.15):
    rng = np.random.Generator(np.random.PCG64(seed))
    lun_g, lun_t = [], []
    for g in range(1, ng + 1):
        for t in (ISTSOIL, ISTCROP, ISTICE, ISTDLAK, ISTWET, URB0, URB0 + 1, URB0 + 2):
            if t == ISTSOIL or rng.random() < 0.35:
                lun_g.append(g); lun_t.append(t)
    lun_g, lun_t = np.array(lun_g, np.int32), np.array(lun_t, np.int32)
    nl = len(lun_t)
    ncol_l = np.where(lun_t >= URB0, 5, np.where(lun_t == ISTCROP, rng.integers(1, 4, nl), 1))
    col_l = np.repeat(np.arange(1, nl + 1, dtype=np.int32), ncol_l)
    nc = len(col_l)
    npat_c = np.where(np.isin(lun_t[col_l - 1], (ISTSOIL,)), rng.integers(1, 16, nc), 1)
    pat_c = np.repeat(np.arange(1, nc + 1, dtype=np.int32), npat_c)
    npch = len(pat_c)
    T = {"lun_itype": lun_t, "lun_lakpoi": (lun_t == ISTDLAK).astype(np.int32), "lun_urbpoi": (lun_t >= URB0).astype(np.int32),
         "lun_active": (rng.random(nl) > p_inactive).astype(np.int32),
         "col_landunit": col_l, "col_gridcell": lun_g[col_l - 1], "col_active": (rng.random(nc) > p_inactive).astype(np.int32),
         "patch_landunit": col_l[pat_c - 1], "patch_active": (rng.random(npch) > p_inactive).astype(np.int32),
         "melt_replaced_by_ice_grc": (rng.random(ng) < 0.4).astype(np.int32)}
    lt_c = lun_t[col_l - 1]
    T["col_hydrologically_active"] = (np.isin(lt_c, (ISTSOIL, ISTCROP)) | ((lt_c >= URB0) & (rng.random(nc) < 0.2))).astype(np.int32)
    lt_p = lun_t[T["patch_landunit"] - 1]
    T["patch_itype"] = np.where(lt_p == ISTCROP, rng.integers(15, 79, npch), np.where(lt_p == ISTSOIL, rng.integers(0, 15, npch), 0)).astype(np.int32)
    b = abi.Bounds()
    b.begg, b.endg, b.begl, b.endl, b.begc, b.endc, b.begp, b.endp = 1, ng, 1, nl, 1, nc, 1, npch
    b.level, b.clump_index = 2, 1
    return b, {k: np.ascontiguousarray(v, dtype=np.int32) for k, v in T.items()}


def expected_lists(b, T, include_inactive, use_cn, use_fates, use_fates_bgc, npcropmin=17, npcropmax=78):
    """numpy statement of every list over the bounds b (1-based indices)."""
    c = np.arange(b.begc, b.endc + 1); p = np.arange(b.begp, b.endp + 1); l = np.arange(b.begl, b.endl + 1)
    ca = (T["col_active"][c - 1] != 0) | bool(include_inactive)
    pa = (T["patch_active"][p - 1] != 0) | bool(include_inactive)
    la = (T["lun_active"][l - 1] != 0) | bool(include_inactive)
    ltc = T["lun_itype"][T["col_landunit"][c - 1] - 1]; ltp = T["lun_itype"][T["patch_landunit"][p - 1] - 1]
    lakc = T["lun_lakpoi"][T["col_landunit"][c - 1] - 1] != 0; lakp = T["lun_lakpoi"][T["patch_landunit"][p - 1] - 1] != 0
    urbc = T["lun_urbpoi"][T["col_landunit"][c - 1] - 1] != 0; urbp = T["lun_urbpoi"][T["patch_landunit"][p - 1] - 1] != 0
    soilc = np.isin(ltc, (ISTSOIL, ISTCROP)); soilp = np.isin(ltp, (ISTSOIL, ISTCROP))
    ivt = T["patch_itype"][p - 1]; crop = (ivt >= npcropmin) & (ivt <= npcropmax)
    melt = T["melt_replaced_by_ice_grc"][T["col_gridcell"][c - 1] - 1] != 0
    nof = not use_fates
    E = {"allc": c[ca], "lakec": c[ca & lakc], "nolakec": c[ca & ~lakc],
         "bgc_soilc": c[ca & soilc] if (use_cn or use_fates_bgc) else c[:0], "soilc": c[ca & soilc],
         "hydrologyc": c[ca & (T["col_hydrologically_active"][c - 1] != 0)], "urbanc": c[ca & urbc], "nourbanc": c[ca & ~urbc],
         "icec": c[ca & (ltc == ISTICE)], "do_smb_c": c[ca & melt & np.isin(ltc, (ISTICE, ISTSOIL))],
         "lakep": p[pa & lakp], "nolakep": p[pa & ~lakp], "nolakeurbanp": p[pa & ~lakp & ~urbp],
         "bgc_vegp": p[pa & soilp] if use_cn else p[:0], "soilp": p[pa & soilp],
         "pcropp": p[pa & crop] if nof else p[:0], "soilnopcropp": p[pa & ~crop & soilp] if nof else p[:0],
         "urbanp": p[pa & urbp], "nourbanp": p[pa & ~urbp], "urbanl": l[la & (T["lun_urbpoi"][l - 1] != 0)],
         "nourbanl": l[la & (T["lun_urbpoi"][l - 1] == 0)]}
    return {k: v.astype(np.int32) for k, v in E.items()}


def make_inputs(alloc, T, include_inactive, use_cn, use_fates, use_fates_bgc, arrays=None):
    fin = abi.FilterInputs()
    fin.alloc = alloc
    src = T if arrays is None else arrays
    for n in abi.FilterInputs._ARRAYS:
        setattr(fin, n, abi.i32p(src[n]))
    fin.include_inactive, fin.use_cn, fin.use_fates, fin.use_fates_bgc = int(include_inactive), int(use_cn), int(use_fates), int(use_fates_bgc)
    fin.npcropmin, fin.npcropmax = 17, 78
    return fin


def make_outputs(b, device=False):
    ext = {"COL": b.endc - b.begc + 1, "PATCH": b.endp - b.begp + 1, "LUN": b.endl - b.begl + 1}
    out = abi.Filters()
    bufs = []
    for k, lev in enumerate(abi.FILTER_LEVEL):
        n = max(ext[lev], 1)
        if device:
            import torch
            a = torch.full((n,), -7, dtype=torch.int32, device="cuda")
        else:
            a = np.full(n, -7, dtype=np.int32)
        bufs.append(a)
        out.list[k] = abi.i32p(a)
    return out, bufs
