"""
PSF fit -> object fit chaining (ngmix/bootstrap.py): the production caller of
the fitters.
"""
from .gexceptions import BootPSFFailure
from .observation import Observation, ObsList, MultiBandObsList


class Bootstrapper(object):
    """
    runner: Runner for the object; psf_runner: PSFRunner, optional;
    ignore_failed_psf: drop observations whose psf fit failed (default True)
    """

    def __init__(self, runner, psf_runner=None, ignore_failed_psf=True):
        self.runner = runner
        self.psf_runner = psf_runner
        self.ignore_failed_psf = ignore_failed_psf

    def go(self, obs):
        return bootstrap(obs=obs, runner=self.runner,
                         psf_runner=self.psf_runner,
                         ignore_failed_psf=self.ignore_failed_psf)

    @property
    def fitter(self):
        return self.runner.fitter


def bootstrap(obs, runner, psf_runner=None, ignore_failed_psf=True):
    """fit the psfs (setting obs.psf.gmix), then the object"""
    if psf_runner is not None:
        psf_runner.go(obs=obs)
        if ignore_failed_psf:
            obs = remove_failed_psf_obs(obs=obs)
    return runner.go(obs=obs)


def _psf_ok(o):
    return o.psf.meta["result"]["flags"] == 0


def remove_failed_psf_obs(obs):
    """observations whose psf fit succeeded, same container type as input;
    BootPSFFailure when a band is left empty"""
    if isinstance(obs, MultiBandObsList):
        out = MultiBandObsList(meta=obs.meta)
        for ol in obs:
            out.append(remove_failed_psf_obs(ol))
        return out
    if isinstance(obs, ObsList):
        out = ObsList(meta=obs.meta)
        for o in obs:
            if _psf_ok(o):
                out.append(o)
        if len(out) == 0:
            raise BootPSFFailure("no good psf fits")
        return out
    if isinstance(obs, Observation):
        if not _psf_ok(obs):
            raise BootPSFFailure("no good psf fits")
        return obs
    raise ValueError('got obs input type: "%s", should be Observation, '
                     'ObsList, or MulitiBandObsList' % type(obs))
