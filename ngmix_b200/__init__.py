"""
ngmix_b200: B200-native batched Gaussian-mixture image fitting, a drop-in for
the fitting hot path of esheldon/ngmix.

The public names mirror ngmix: GMix/GMixModel, Observation/Jacobian/
make_pixels, fitting.Fitter (Levenberg-Marquardt), em and admom runners.
Every numerical operation goes through the C ABI (include/ngx.h) to
hand-written sm_100a CUDA kernels; there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import flags, defaults, gexceptions, priors, shape  # noqa: F401
from .gexceptions import GMixRangeError, GMixFatalError  # noqa: F401
from .jacobian import Jacobian, DiagonalJacobian, UnitJacobian  # noqa: F401
from .pixels import make_pixels, make_coords  # noqa: F401
from .gmix import (  # noqa: F401
    GMix, GMixModel, GMixCM, GMixCoellip, make_gmix_model, gmix_concat,
)
from .observation import (  # noqa: F401
    Observation, ObsList, MultiBandObsList, get_mb_obs,
)
from . import fitting, em, admom  # noqa: F401
from .fitting import (  # noqa: F401
    Fitter, CoellipFitter, PSFFluxFitter, fit_batch, fit_batch_multi,
    fit_batch_coellip,
)
from . import guessers, runners, bootstrap  # noqa: F401
from .em import EMFitter, run_em, em_batch  # noqa: F401
from .admom import AdmomFitter, run_admom, admom_batch  # noqa: F401
from . import joint_prior, moments, gaussmom, gaussap  # noqa: F401
from .gaussmom import GaussMom, gaussmom_batch  # noqa: F401
from .batch import ObservationBatch  # noqa: F401
from ._lib import pinned, pinned_empty, pinned_copy  # noqa: F401
