"""
Run a fitter with guesses from a guesser, retrying on failure
(ngmix/runners.py).  Host-side control flow around the device fitters: the
fits themselves are ngx_lm_fit / ngx_em / ngx_admom launches.
"""
from .observation import Observation, ObsList, MultiBandObsList


class RunnerBase(object):
    """
    fitter: object with go(obs=, guess=) (or go(obs=) without a guesser)
    guesser: callable guesser(obs=obs) -> guess, optional
    ntry: number of attempts while flags != 0
    """

    def __init__(self, fitter, guesser=None, ntry=1):
        self.fitter = fitter
        self.guesser = guesser
        self.ntry = ntry


class Runner(RunnerBase):
    """run the fitter on the observation(s) themselves"""

    def go(self, obs):
        return run_fitter(obs=obs, fitter=self.fitter, guesser=self.guesser,
                          ntry=self.ntry)


class PSFRunner(RunnerBase):
    """run the fitter on the psf observation of every observation; with
    set_result the result goes to obs.psf.meta['result'] and a successful
    fit's mixture to obs.psf.gmix"""

    def __init__(self, fitter, guesser=None, ntry=1, set_result=True):
        super().__init__(fitter, guesser=guesser, ntry=ntry)
        self.set_result = set_result

    def go(self, obs):
        return run_psf_fitter(obs=obs, fitter=self.fitter,
                              guesser=self.guesser, ntry=self.ntry,
                              set_result=self.set_result)


def run_fitter(obs, fitter, guesser=None, ntry=1):
    """up to ntry fits, each from a fresh guess; stops at flags == 0"""
    res = None
    for _ in range(ntry):
        if guesser is None:
            res = fitter.go(obs=obs)
        else:
            res = fitter.go(obs=obs, guess=guesser(obs=obs))
        if res["flags"] == 0:
            break
    return res


def run_psf_fitter(obs, fitter, guesser=None, ntry=1, set_result=True):
    """
    Observation -> result; ObsList -> list of results; MultiBandObsList ->
    list of lists.  The psf observation is fit when there is one, else the
    observation itself.
    """
    if isinstance(obs, MultiBandObsList):
        return [run_psf_fitter(obs=ol, fitter=fitter, guesser=guesser,
                               ntry=ntry, set_result=set_result)
                for ol in obs]
    if isinstance(obs, ObsList):
        return [run_psf_fitter(obs=o, fitter=fitter, guesser=guesser,
                               ntry=ntry, set_result=set_result)
                for o in obs]
    if not isinstance(obs, Observation):
        raise ValueError(
            "obs must be an Observation, ObsList, or MultiBandObsList")
    target = obs.psf if obs.has_psf() else obs
    res = run_fitter(obs=target, fitter=fitter, guesser=guesser, ntry=ntry)
    if set_result:
        target.meta["result"] = res
        if res["flags"] == 0 and hasattr(res, "get_gmix"):
            target.gmix = res.get_gmix()
    return res


def run_fit_batch(model, images, weights, jacobians, psf_pars, guesser,
                  ntry=1, **kw):
    """
    The Runner's retry loop (run_fitter, ngmix/runners.py:116-147) over a
    batch: round 0 fits every object, each later round refits only the
    objects whose flags are still non-zero, from fresh guesses, in one launch.

    guesser: callable(indices, itry) -> [len(indices), npars] guesses for the
        objects `indices` (int array) in round `itry`
    other arguments as fit_batch.  Returns (records [n], tries [n]) where
    tries counts the fits each object went through.
    """
    import numpy as np
    from .fitting import fit_batch
    images, weights = np.asarray(images), np.asarray(weights)
    n = images.shape[0]
    jac = np.asarray(jacobians).reshape(n, -1)
    psf = None if psf_pars is None else np.asarray(psf_pars)
    noise = kw.pop("noise_images", None)
    todo = np.arange(n)
    recs = None
    tries = np.zeros(n, dtype="i4")
    for itry in range(ntry):
        guess = np.asarray(guesser(todo, itry), dtype="f8")
        sub = fit_batch(model, images[todo], weights[todo], jac[todo],
                        None if psf is None else psf[todo], guess,
                        noise_images=None if noise is None
                        else np.asarray(noise)[todo], **kw)
        if recs is None:
            recs = sub.copy()
        else:
            recs[todo] = sub
        tries[todo] += 1
        todo = todo[sub["flags"] != 0]
        if todo.size == 0:
            break
    return recs, tries
