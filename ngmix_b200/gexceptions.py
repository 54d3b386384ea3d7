"""exception classes of the public API (ngmix/gexceptions.py)"""


class NGmixBaseException(Exception):
    def __init__(self, value):
        super().__init__(value)
        self.value = value

    def __str__(self):
        return repr(self.value)


class GMixRangeError(NGmixBaseException):
    """Some number was out of range."""


class GMixFatalError(NGmixBaseException):
    """A fatal error in the Gaussian mixtures."""


class GMixMaxIterEM(NGmixBaseException):
    """EM algorithm hit max iter."""


class PSFFluxFailure(NGmixBaseException):
    """Failure to fit PSF fluxes."""


class BootPSFFailure(NGmixBaseException):
    """Failure to bootstrap PSF"""


class BootGalFailure(NGmixBaseException):
    """Failure to bootstrap galaxy"""
