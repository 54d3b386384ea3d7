import time, numpy as np, torch, sys
sys.path.insert(0,'.')
from ngmix_b200 import _lib
import bench
_lib.init(0)
n=10000
x=torch.empty(n*2304, dtype=torch.float64).pin_memory()
d=torch.empty_like(x, device='cuda')
for _ in range(3): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(5): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
dt=(time.perf_counter()-t0)/5
print("H2D pinned %.1f MB in %.2f ms -> %.1f GB/s" % (x.numel()*8/1e6, dt*1e3, x.numel()*8/dt/1e9))
b=bench.make_inputs(n, 0)
pin=lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
hi,hw,hj,hp,hg=[pin(b[k]) for k in ("images","weights","jac","psf","guess")]
hr=torch.zeros(n*_lib.LM_RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
out=np.frombuffer(hr.numpy(), dtype=_lib.LM_RESULT_DTYPE)
for name,kw in (("pageable out",{}),("pinned out",{"out":out})):
    for _ in range(2): _lib.lm_fit("exp", hi.numpy(),hw.numpy(),hj.numpy(),hp.numpy(),hg.numpy(),**kw)
    t0=time.perf_counter()
    for _ in range(5): _lib.lm_fit("exp", hi.numpy(),hw.numpy(),hj.numpy(),hp.numpy(),hg.numpy(),**kw)
    dt=(time.perf_counter()-t0)/5
    print(name, "%.2f ms/step -> %.0f fits/s"%(dt*1e3, n/dt))
