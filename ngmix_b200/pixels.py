"""
make_pixels / make_coords: mirror of ngmix/pixels/pixels.py.  The arrays have
the reference's _pixels_dtype / _coords_dtype layout and are built on the
device (ngx_make_pixels / ngx_make_coords: v,u from the jacobian,
ierr = sqrt(weight), ordered compaction of positive-weight pixels).
"""
import numpy as np

from . import _lib
from .gexceptions import GMixFatalError

__all__ = ['make_pixels', 'make_coords']

_pixels_dtype = _lib.PIXEL_DTYPE
_coords_dtype = _lib.COORD_DTYPE


def make_pixels(image, weight, jacob, ignore_zero_weight=True):
    """
    1-d pixel array (u, v, area, val, ierr, fdiff) for an image and weight.

    With ignore_zero_weight (default) pixels of weight <= 0 are dropped, so
    the array length is the number of positive weights; GMixFatalError is
    raised if there are none (pixels.py:33-37).
    """
    image = np.asarray(image, dtype="f8")
    weight = np.asarray(weight, dtype="f8")
    if ignore_zero_weight and not np.any(weight > 0.0):
        raise GMixFatalError("no weights > 0")
    pixels, npix = _lib.make_pixels(image, weight, jacob._data,
                                    ignore_zero_weight=ignore_zero_weight)
    return pixels[0, :npix[0]].copy()


def make_coords(dims, jacob):
    """1-d coords array (u, v, area) for every pixel of a dims image"""
    nrow, ncol = dims
    return _lib.make_coords((int(nrow), int(ncol)), jacob._data)[0].copy()
