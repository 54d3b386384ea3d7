"""k_admom: rate on 100 000 PSF stamps and parity against the oracle"""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
from ngmix_b200 import _lib, sims_device
import oracle
_lib.init(0)
dev = torch.device("cuda", 0)
n = 100000
t = sims_device.make_psf_batch_device(n, 3000000, dev)
d_am = torch.zeros(n * _lib.ADMOM_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
def run():
    _lib.admom_device(n, t["images"].data_ptr(), t["weights"].data_ptr(),
                      t["jac"].data_ptr(), 25, 25, t["am_guess"].data_ptr(), d_am.data_ptr())
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
ram = np.frombuffer(d_am.cpu().numpy().tobytes(), dtype=_lib.ADMOM_RESULT_DTYPE)
im, wt = t["images"].cpu().numpy(), t["weights"].cpu().numpy()
oam = oracle.admom_go_batch(im, wt, t["h_jac"], t["h_am_guess"], nthreads=os.cpu_count())
good = (oam["flags"] == 0) & (ram["flags"] == 0)
rel = np.abs(ram["pars"][good] - oam["pars"][good]) / np.maximum(np.abs(oam["pars"][good]), 1e-12)
out = {"n": n, "ms": ms, "fits_per_s": n / ms * 1e3,
       "numiter_differ": int((oam["numiter"] != ram["numiter"]).sum()),
       "flags_differ": int((oam["flags"] != ram["flags"]).sum()),
       "max_rel_dev_T_flux": float(rel[:, 4:6].max()),
       "max_rel_dev_sums": float(np.max(np.abs(ram["sums"][good] - oam["sums"][good]) / np.maximum(np.abs(oam["sums"][good]), 1e-300)))}
print(json.dumps(out), flush=True)
