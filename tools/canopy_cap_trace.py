"""tools/canopy_cap_trace.py: which exposed-vegetation patches use all 41 ITERATION passes of CanopyFluxes, and why:
statistics from the oracle, per-pass traces of three of them from the independent Python restatement (tests/canopy_python.py). CPU only."""
import sys; sys.path.insert(0,'/root/repo')
import numpy as np, ctypes as C
from types import SimpleNamespace
from ctsm_b200 import abi, synthetic_canopy
from tests import canopy_python as cp, phs_python as pp
from tests.test_oracle_canopy_pin import canopy_patch_inputs
from tests.util import copy_state
from oracle import oracle
sg,S=synthetic_canopy.make_full_case(3000,seed=77)
prm=abi.default_params()
S0=copy_state(S)
L=oracle.lib()
fe=sg.filters["exposedvegp"]
f=abi.make_struct("canopyfluxes",S,sg.bounds); st=abi.Status()
L.oracle_canopyfluxes(C.byref(prm),C.byref(sg.bounds),len(fe),abi.i32p(fe),C.byref(f),C.byref(st))
it=S["num_iter"][fe-1]
cap=fe[it>40]-1
print("exposed",len(fe),"capped",len(cap), "hist", np.bincount(it.astype(int))[:45])
c=S0["column"][cap]-1; g=S0["gridcell"][cap]-1
allc=S0["column"][fe-1]-1; allg=S0["gridcell"][fe-1]-1
def stat(name,a_cap,a_all): print("%-14s capped mean %.4g  all mean %.4g"%(name,np.mean(a_cap),np.mean(a_all)))
wind=lambda gi: np.hypot(S0["forc_u"][gi],S0["forc_v"][gi])
stat("wind",wind(g),wind(allg))
stat("night",(S0["parsun_z"][0,cap]<=0),(S0["parsun_z"][0,fe-1]<=0))
stat("thm-t_grnd",S0["thm"][cap]-S0["t_grnd"][c],S0["thm"][fe-1]-S0["t_grnd"][allc])
stat("elai+esai",S0["elai"][cap]+S0["esai"][cap],S0["elai"][fe-1]+S0["esai"][fe-1])
stat("htop",S0["htop"][cap],S0["htop"][fe-1])
stat("zeta_final",S["zeta"][cap],S["zeta"][fe-1])
stat("sabv",S0["sabv"][cap],S0["sabv"][fe-1])
stat("snow_depth",S0["snow_depth"][c],S0["snow_depth"][allc])
stat("fwet",S0["fwet"][cap],S0["fwet"][fe-1])
stat("h2ocan",S0["liqcan"][cap]+S0["snocan"][cap],S0["liqcan"][fe-1]+S0["snocan"][fe-1])
# trace three capped patches by instrumenting the restatement
import math
M=SimpleNamespace(**{k:getattr(prm,k) for k,_ in abi.Params._fields_ if not k.startswith("reserved")})
orig_fv=cp.friction_velocity
for p in cap[:3]:
    P=canopy_patch_inputs(S0,int(p)); P.gs_mol_patch=0.0
    log=[]
    def fvwrap(*a):
        o=orig_fv(*a); log.append((a[8], a[7], a[10], o.ustar)); return o
    cp.friction_velocity=fvwrap
    O=cp.canopy_fluxes_patch(P,M,pp.photosynthesis_hydraulic_stress)
    cp.friction_velocity=orig_fv
    print("patch",p,"night",P.par_z[1]<=0,"wind",math.hypot(P.forc_u,P.forc_v),"iters",O.num_iter)
    for (it,obu,um,ustar) in log[:12]+log[-6:]: print("   it %2d obu %12.4f um %8.4f ustar %8.5f zeta %8.4f"%(it,obu,um,ustar,(O.forc_hgt_u_patch-O.displa)/obu))
w_cap=wind(g); w_all=wind(allg)
print("capped with wind<2.5: %.2f ; all with wind<2.5: %.3f ; cap rate | wind<2.5: %.4f ; cap rate | wind>=2.5: %.5f"%((w_cap<2.5).mean(),(w_all<2.5).mean(), (w_cap<2.5).sum()/max((w_all<2.5).sum(),1), (w_cap>=2.5).sum()/max((w_all>=2.5).sum(),1)))
print("wind quantiles all", np.quantile(w_all,[0.05,0.25,0.5,0.75,0.95]))
