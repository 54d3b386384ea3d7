import sys, numpy as np, torch
sys.path.insert(0,'.')
from ngmix_b200 import _lib, sims
import bench
_lib.init(0)
n=10000
peak,_=_lib.peak_dfma(60000)
b=bench.make_inputs(n,0)
gm0,_=_lib.gmix_fill("exp", b["guess"]); gp,_=_lib.gmix_fill("full", b["psf"].reshape(n,-1))
gm=_lib.gmix_convolve(gm0,gp); _lib.gmix_set_norms(gm)
dev=torch.device("cuda",0)
d_gm=torch.from_numpy(np.frombuffer(gm.tobytes(),dtype=np.uint8).copy()).to(dev)
d_im=torch.from_numpy(b["images"]).to(dev); d_wt=torch.from_numpy(b["weights"]).to(dev); d_j=torch.from_numpy(b["jac"]).to(dev)
d_o=torch.zeros(n*4,dtype=torch.float64,device=dev); d_f=torch.zeros(n*2304,dtype=torch.float64,device=dev)
s=torch.cuda.current_stream().cuda_stream
for name,fd in (("loglike only",0),("with fdiff",d_f.data_ptr())):
    f=lambda: _lib.loglike_device(n,18,d_gm.data_ptr(),d_im.data_ptr(),d_wt.data_ptr(),d_j.data_ptr(),48,48,d_o.data_ptr(),d_fdiff=fd,stream=s)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    evals=n*2304*18
    print("%s: %.3f ms, %.3e gauss-pixel evals/s, %.3e model-pixel evals/s, fp64-instr frac %.3f, flops frac %.3f" % (name, ms, evals/ms*1e3, n*2304/ms*1e3, evals*16/ms*1e3/peak, evals*28/ms*1e3/(2*peak)))
