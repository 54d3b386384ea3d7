#!/usr/bin/env python
"""k_admom alone, for `ncu --set full -k regex:k_admom -c 1` (the multi-kernel
capture returned no counters for this launch)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ngmix_b200 import _lib, sims_device
_lib.init(0)
dev = torch.device("cuda", 0)
n = 20000
p = sims_device.make_psf_batch_device(n, 1, dev)
d_am = torch.zeros(n * _lib.ADMOM_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.admom_device(n, p["images"].data_ptr(), p["weights"].data_ptr(),
                      p["jac"].data_ptr(), 25, 25, p["am_guess"].data_ptr(),
                      d_am.data_ptr(), stream=s)
torch.cuda.synchronize()
