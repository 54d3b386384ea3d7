"""joint priors under the reference's module name (ngmix/joint_prior.py)"""
from .priors import (  # noqa: F401
    PriorSimpleSep, PriorBDFSep, PriorBDSep, PriorCoellipSame)
