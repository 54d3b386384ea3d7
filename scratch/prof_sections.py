import sys, numpy as np
sys.path.insert(0,'.')
from ngmix_b200 import _lib
_lib.LIBPATH = 'scratch/libngx_prof.so'
import bench
_lib.init(0)
n=10000
b=bench.make_inputs(n,0)
for _ in range(2):
    res=_lib.lm_fit("exp", b["images"],b["weights"],b["jac"],b["psf"],b["guess"])
t=res["pars_cov0"][:,56:61]
tot=t.sum(axis=1)
names=["stage","mixture","pass+reduce","solve(+barrier)","finish"]
print("mean cycles per fit per CTA: total %.0f"%tot.mean())
for k,nm in enumerate(names): print("  %-16s %10.0f  %5.1f%%"%(nm,t[:,k].mean(),100*t[:,k].mean()/tot.mean()))
print("nfev mean",res["nfev"].mean(),"njev",res["njev"].mean())
