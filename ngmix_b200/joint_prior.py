"""joint priors under the reference's module name (ngmix/joint_prior.py)"""
from .priors import PriorSimpleSep, PriorBDFSep, PriorBDSep  # noqa: F401
