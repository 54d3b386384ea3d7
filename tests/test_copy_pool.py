"""
The host threads that stage pageable arrays into pinned slots
(ngmix_b200/csrc/ngx_copy_pool.h): a plain copy for every size and alignment,
with the default pool and with one thread.  No GPU needed.
"""
import ctypes as C
import os
import subprocess
import sys

import pytest

from conftest import ROOT

HDIR = os.path.join(ROOT, "tests", "host_harness")


@pytest.fixture(scope="module")
def so():
    out = os.path.join(HDIR, "libcopypool.so")
    src = os.path.join(HDIR, "copy_pool_harness.cpp")
    hdr = os.path.join(ROOT, "ngmix_b200", "csrc", "ngx_copy_pool.h")
    if (not os.path.exists(out)
            or os.path.getmtime(out) < max(os.path.getmtime(src),
                                           os.path.getmtime(hdr))):
        subprocess.check_call([
            "g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread",
            "-I/usr/local/cuda/include", "-o", out, src])
    return out


@pytest.mark.parametrize("threads", ["", "1", "5"])
def test_copy_pool_copies_exactly(so, threads):
    # NGX_STAGE_THREADS is read when the process-wide pool is created
    code = ("import ctypes as C, sys; "
            "sys.exit(C.CDLL(%r).copy_pool_check(%d))"
            % (so, int(threads) if threads else 0))
    env = dict(os.environ)
    if threads:
        env["NGX_STAGE_THREADS"] = threads
    else:
        env.pop("NGX_STAGE_THREADS", None)
    assert subprocess.call([sys.executable, "-c", code], env=env) == 0
