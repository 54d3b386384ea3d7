"""
TEST INFRASTRUCTURE - the north-star CPU arm: the UNMODIFIED reference
(ngmix: numba kernels + scipy MINPACK) run through its own public API
(`ngmix.fitting.Fitter.go`, `ngmix.em.EMFitter.go`, `ngmix.admom.AdmomFitter.go`),
one process per core with `multiprocessing` (fork), objects chunked by index
("CPU baseline timing").

The reference package is staged by oracle/make_ref.sh into oracle/_ref/
(git-ignored, shipped to the GPU box by gpurun).  Only bench.py's
`--impl reference` / `cpu_baseline` legs and the tests import this module;
nothing under ngmix_b200/ does.

Timing protocol: the parent imports ngmix, builds every `Observation`
(untimed, reported separately) and runs one warm-up fit (numba JIT) BEFORE
forking, so the workers inherit the objects and the compiled code; a step
runs only the `fitter.go` loops, and its wall time is measured in the parent
around the whole pool call (slowest worker + dispatch).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")

_ngmix = None
_state = {}


def available():
    """(ok, why): the staged reference is importable here"""
    if not os.path.isdir(os.path.join(REFDIR, "ngmix")):
        return False, "oracle/_ref/ngmix not staged (run oracle/make_ref.sh)"
    try:
        import numba  # noqa: F401
        import scipy  # noqa: F401
    except ImportError as e:
        return False, "missing dependency: %s" % e
    return True, ""


def import_ref():
    """import the staged reference (never /root/reference at run time)"""
    global _ngmix
    if _ngmix is None:
        os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/ngx_numba_cache")
        os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
        if REFDIR not in sys.path:
            sys.path.insert(0, REFDIR)
        import ngmix
        assert os.path.realpath(ngmix.__file__).startswith(
            os.path.realpath(REFDIR)), ngmix.__file__
        _ngmix = ngmix
    return _ngmix


def _jacobian(ngmix, rec):
    return ngmix.Jacobian(row=rec[0], col=rec[1], dvdrow=rec[2],
                          dvdcol=rec[3], dudrow=rec[4], dudcol=rec[5])


def _make_obs(ngmix, b, i, with_psf=True):
    jac = _jacobian(ngmix, b["jac"][i])
    psf_obs = None
    if with_psf and b.get("psf") is not None:
        gm = ngmix.GMix(pars=np.asarray(b["psf"][i]).ravel())
        # the psf image itself is never read by Fitter.go once gmix is set
        psf_obs = ngmix.Observation(np.zeros((5, 5)) + 1.0,
                                    jacobian=ngmix.DiagonalJacobian(
                                        row=2, col=2, scale=float(
                                            b["jac"][i][7])),
                                    gmix=gm)
    return ngmix.Observation(b["images"][i], weight=b["weights"][i],
                             jacobian=jac, psf=psf_obs)


def make_prior(ngmix, kind, model="exp"):
    """the prior sets of SURVEY 8d config 2 / config 4"""
    if kind in (None, "none"):
        return None
    pr = ngmix.priors
    cen = pr.CenPrior(0.0, 0.0, 0.263, 0.263)
    g = pr.GPriorBA(0.3)
    if model == "bdf":
        return ngmix.joint_prior.PriorBDFSep(
            cen_prior=cen, g_prior=g, T_prior=pr.FlatPrior(-1.0, 1.0e3),
            fracdev_prior=pr.Normal(0.5, 0.1),
            F_prior=pr.FlatPrior(-100.0, 1.0e9))
    if kind == "flat":
        return ngmix.joint_prior.PriorSimpleSep(
            cen_prior=cen, g_prior=g, T_prior=pr.FlatPrior(-1.0, 1.0e3),
            F_prior=pr.FlatPrior(-100.0, 1.0e9))
    return ngmix.joint_prior.PriorSimpleSep(
        cen_prior=cen, g_prior=g,
        T_prior=pr.TwoSidedErf(-1.0, 0.1, 1.0e3, 1.0),
        F_prior=pr.TwoSidedErf(-1.0e2, 1.0, 1.0e9, 1.0e8))


# ---------------------------------------------------------------------------
# worker side (forked: _state, with the Observation objects, is inherited
# copy-on-write from the parent)

def _work(task):
    kind, lo, hi = task
    ngmix = _ngmix
    b = _state["batch"]
    obs = _state["obs"]
    t0 = time.perf_counter()
    flags = np.zeros(hi - lo, dtype="i4")
    if kind == "lm":
        fitter = ngmix.fitting.Fitter(model=_state["model"],
                                      prior=_state["prior"])
        pars = np.zeros((hi - lo, b["guess"].shape[1]))
        nfev = np.zeros(hi - lo, dtype="i4")
        for k in range(hi - lo):
            r = fitter.go(obs=obs[lo + k], guess=b["guess"][lo + k].copy())
            flags[k] = r["flags"]
            nfev[k] = r["nfev"]
            pars[k] = r["pars"]
        out = dict(flags=flags, nfev=nfev, pars=pars)
    elif kind in ("em", "admom"):
        fitter = ngmix.em.EMFitter(tol=1.0e-5) if kind == "em" \
            else ngmix.admom.AdmomFitter()
        gkey = "em_guess" if kind == "em" else "am_guess"
        numiter = np.zeros(hi - lo, dtype="i4")
        for k in range(hi - lo):
            guess = ngmix.GMix(pars=np.asarray(b[gkey][lo + k]).ravel())
            r = fitter.go(obs=obs[lo + k], guess=guess)
            numiter[k] = r["numiter"]
            flags[k] = r["flags"]
        out = dict(flags=flags, numiter=numiter)
    else:
        raise ValueError(kind)
    out["t_fit"] = time.perf_counter() - t0
    out["lo"] = lo
    return out


def time_batch(b, kind="lm", model="exp", prior_kind=None, nproc=None,
               steps=3, warmup=1, max_seconds=150.0, chunks_per_proc=4):
    """
    Fit the batch `b` with the reference on nproc processes.

    Returns dict(value = fits/s over the timed steps, ms_per_step, steps,
    cores, n, per_core_fits_per_s = from the workers' own fit-loop clocks,
    obs_construction_s_per_object, results = concatenated outputs of the last
    step).
    """
    import multiprocessing as mp
    ngmix = import_ref()
    nproc = nproc or os.cpu_count()
    n = b["images"].shape[0]
    prior = make_prior(ngmix, prior_kind, model) if kind == "lm" else None
    t0 = time.perf_counter()
    obs = [_make_obs(ngmix, b, i, with_psf=(kind == "lm")) for i in range(n)]
    t_obs = time.perf_counter() - t0
    _state.update(batch=b, model=model, prior=prior, obs=obs)
    # one fit in the parent compiles every numba kernel on the path, so the
    # forked workers inherit the machine code
    _work((kind, 0, min(2, n)))
    # contiguous index ranges handed out dynamically (Pool.map, chunksize 1)
    nchunk = max(1, min(n, nproc * chunks_per_proc))
    edges = np.linspace(0, n, nchunk + 1).astype(int)
    tasks = [(kind, int(edges[i]), int(edges[i + 1])) for i in range(nchunk)
             if edges[i + 1] > edges[i]]
    ctx = mp.get_context("fork")
    times, last = [], None
    with ctx.Pool(nproc) as pool:
        for step in range(warmup + steps):
            t0 = time.perf_counter()
            outs = pool.map(_work, tasks, chunksize=1)
            dt = time.perf_counter() - t0
            if step >= warmup:
                times.append(dt)
                last = outs
                if sum(times) > max_seconds:
                    break
    last = sorted(last, key=lambda o: o["lo"])
    res = {k: np.concatenate([o[k] for o in last])
           for k in last[0] if k not in ("t_fit", "lo")}
    ms = 1e3 * float(np.mean(times))
    fit_core_s = float(np.sum([o["t_fit"] for o in last]))
    _state.pop("obs", None)
    return dict(value=n / (ms * 1e-3), ms_per_step=ms, steps=len(times),
                cores=nproc, n=n,
                per_core_fits_per_s=n / fit_core_s,
                obs_construction_s_per_object=t_obs / max(n, 1),
                results=res)


def single_fit_latency(b, model="exp", nrep=20):
    """config 1: wall time of one Fitter.go on one stamp (JIT warm), median
    over nrep repeats; Observation construction excluded"""
    ngmix = import_ref()
    obs = _make_obs(ngmix, b, 0)
    fitter = ngmix.fitting.Fitter(model=model)
    fitter.go(obs=obs, guess=b["guess"][0].copy())
    ts = []
    for _ in range(nrep):
        t0 = time.perf_counter()
        r = fitter.go(obs=obs, guess=b["guess"][0].copy())
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), r
