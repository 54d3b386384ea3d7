"""
CPU check of the engine's on-device LM state machine
(ngmix_b200/csrc/ngx_lm_solve.cuh compiled as plain C++): fed only
(J^T J, J^T f, |f|) - computed here by the oracle - it must reproduce the
reference's MINPACK results in the golden fixtures (ier, nfev, parameters).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import load_golden, ROOT

HDIR = os.path.join(ROOT, "tests", "host_harness")


@pytest.fixture(scope="module")
def harness(ora):
    so = os.path.join(HDIR, "libharness.so")
    src = os.path.join(HDIR, "lm_solve_harness.cpp")
    hdr = os.path.join(ROOT, "ngmix_b200", "csrc", "ngx_lm_solve.cuh")
    stale = (not os.path.exists(so)
             or os.path.getmtime(so) < max(os.path.getmtime(src),
                                           os.path.getmtime(hdr),
                                           os.path.getmtime(ora.LIBPATH)))
    if stale:
        subprocess.check_call([
            "g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", so,
            src, "-L" + os.path.join(ROOT, "oracle"), "-loracle",
            "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    ora.lib()
    return C.CDLL(so)


def _fit(harness, ora, model, image, weight, jac, psf, guess, prior):
    pix = ora.make_pixels(image, weight, jac)
    psf = np.ascontiguousarray(psf, dtype="f8")
    guess = np.ascontiguousarray(guess, dtype="f8")
    pars = np.zeros(8)
    cov0 = np.zeros(64)
    nfev = C.c_int()
    ier = C.c_int()
    have = C.c_int()
    harness.harness_lm_fit(
        C.c_int(ora.MODELS[model]), pix.ctypes.data_as(C.c_void_p),
        C.c_int(pix.size), psf.ctypes.data_as(C.c_void_p),
        C.c_int(psf.size // 6), C.byref(prior),
        guess.ctypes.data_as(C.c_void_p), C.c_int(4000), C.c_double(1e-5),
        C.c_double(1e-5), pars.ctypes.data_as(C.c_void_p), C.byref(nfev),
        C.byref(ier), cov0.ctypes.data_as(C.c_void_p), C.byref(have))
    n = guess.size
    return pars[:n], nfev.value, ier.value, cov0[:n * n].reshape(n, n), \
        have.value


@pytest.mark.parametrize("fname,tag,pk,model", [
    ("lm_exp", "c1", "none", "exp"), ("lm_exp", "hi", "none", "exp"),
    ("lm_exp", "hi", "flat", "exp"), ("lm_exp", "lo", "none", "exp"),
    ("lm_exp", "lo", "flat", "exp"), ("lm_exp", "lo", "erf", "exp"),
    ("lm_other", "gauss", "none", "gauss"), ("lm_other", "dev", "flat", "dev"),
    ("lm_other", "bad", "none", "exp"), ("lm_other", "bad", "flat", "exp"),
])
def test_normal_equation_lm_matches_minpack(harness, ora, fname, tag, pk,
                                            model):
    g = load_golden(fname)
    pre = "%s_%s_" % (tag, pk)
    prior = ora.make_prior(pk)
    n = g[tag + "_images"].shape[0]
    nbad = 0
    worst = 0.0
    for i in range(n):
        pars, nfev, ier, cov0, have = _fit(
            harness, ora, model, g[tag + "_images"][i],
            g[tag + "_weights"][i], g[tag + "_jac"][i], g[tag + "_psf"][i],
            g[tag + "_guess"][i], prior)
        if nfev != g[pre + "nfev"][i] or ier != g[pre + "ier"][i]:
            nbad += 1
            continue
        ref = g[pre + "pars"][i]
        # north_star tolerance for fitted parameters is 1e-5 relative
        err = np.max(np.abs(pars - ref) / np.maximum(np.abs(ref), 1e-3))
        worst = max(worst, err)
        assert have == 1
        c_ref = g[pre + "pars_cov0"][i]
        sc = np.sqrt(np.outer(np.diag(c_ref), np.diag(c_ref)))
        assert np.max(np.abs(cov0 - c_ref) / sc) < 1e-5
    # squaring the condition number costs a few digits at low S/N; the
    # contract (north_star) is 1e-5
    assert worst < 1e-6, worst
    assert nbad == 0, "%d of %d differ in nfev/ier" % (nbad, n)


def test_cov_flags(harness):
    a = np.diag([1.0, 2.0, 3.0])
    f = harness.harness_test_cov
    assert f(3, a.ctypes.data_as(C.c_void_p)) == 0
    b = a.copy()
    b[1, 1] = -1.0
    assert f(3, b.ctypes.data_as(C.c_void_p)) == 1024 | 2048
    c = np.array([[1.0, 2.0], [2.0, 1.0]])       # indefinite
    assert f(2, c.ctypes.data_as(C.c_void_p)) == 1024
    d = a.copy()
    d[0, 2] = np.nan
    assert f(3, d.ctypes.data_as(C.c_void_p)) == 8192
