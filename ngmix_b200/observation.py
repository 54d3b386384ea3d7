"""
Observation / ObsList / MultiBandObsList: the real-space part of
ngmix/observation.py with the same constructor, setters and properties.
The per-pixel array `obs.pixels` is built by the device (pixels.make_pixels)
whenever the image, weight or jacobian change, as the reference rebuilds it
with numba (observation.py:814-830).  The k-space classes need galsim and are
out of scope.
"""
import copy

import numpy as np

from .gmix import GMix
from .jacobian import Jacobian, UnitJacobian
from .pixels import make_pixels

__all__ = ['Observation', 'ObsList', 'MultiBandObsList', 'get_mb_obs']


class MetadataMixin(object):
    @property
    def meta(self):
        return self._meta

    @meta.setter
    def meta(self, meta):
        self.set_meta(meta)

    def set_meta(self, meta):
        if meta is None:
            meta = {}
        if not isinstance(meta, dict):
            raise TypeError("meta data must be in dictionary form, got %s"
                            % type(meta))
        self._meta = meta

    def update_meta_data(self, meta):
        if not isinstance(meta, dict):
            raise TypeError("meta data must be in dictionary form, got %s"
                            % type(meta))
        self.meta.update(meta)


def _optional_array(name):
    """property trio has_X / set_X / X for an optional image-shaped array"""
    attr = '_' + name

    def has(self):
        return hasattr(self, attr)

    def setter(self, value):
        if value is None:
            if hasattr(self, attr):
                delattr(self, attr)
            return
        value = np.asarray(value)
        assert len(value.shape) == 2, "%s must be 2d" % name
        assert value.shape == self.image.shape, \
            "image and %s must be same shape" % name
        setattr(self, attr, value)

    def getter(self):
        return self._get_view(getattr(self, attr))

    return has, setter, property(getter, lambda self, v: setter(self, v))


class Observation(MetadataMixin):
    """
    An image with its weight map, jacobian and optional psf Observation,
    gmix, masks and noise image.
    """

    def __init__(self, image, weight=None, bmask=None, ormask=None,
                 noise=None, jacobian=None, gmix=None, psf=None, meta=None,
                 mfrac=None, store_pixels=True, ignore_zero_weight=True):
        self._writeable = False
        self._ignore_zero_weight = ignore_zero_weight
        self._store_pixels = store_pixels
        self.set_image(image, update_pixels=False)
        self.set_weight(weight, update_pixels=False)
        self.set_jacobian(jacobian, update_pixels=False)
        self.update_pixels()
        self.set_meta(meta)
        self.set_bmask(bmask)
        self.set_ormask(ormask)
        self.set_noise(noise)
        self.set_gmix(gmix)
        self.set_psf(psf)
        self.set_mfrac(mfrac)

    has_bmask, set_bmask, bmask = _optional_array('bmask')
    has_ormask, set_ormask, ormask = _optional_array('ormask')
    has_noise, set_noise, noise = _optional_array('noise')
    has_mfrac, set_mfrac, mfrac = _optional_array('mfrac')

    @property
    def image(self):
        return self._get_view(self._image)

    @image.setter
    def image(self, image):
        self.set_image(image)

    @property
    def weight(self):
        return self._get_view(self._weight)

    @weight.setter
    def weight(self, weight):
        self.set_weight(weight)

    @property
    def pixels(self):
        return self._pixels

    @property
    def jacobian(self):
        return self.get_jacobian()

    @jacobian.setter
    def jacobian(self, jacobian):
        self.set_jacobian(jacobian)

    @property
    def gmix(self):
        return self.get_gmix()

    @gmix.setter
    def gmix(self, gmix):
        self.set_gmix(gmix)

    @property
    def psf(self):
        return self._psf

    @psf.setter
    def psf(self, psf):
        self.set_psf(psf)

    def set_image(self, image, update_pixels=True):
        image_old = getattr(self, '_image', None)
        image = np.asarray(image, dtype='f8')
        assert len(image.shape) == 2, "image must be 2d"
        if image_old is not None:
            assert image.shape == image_old.shape, (
                "old and new image must have same shape, to maintain "
                "consistency, got %s vs %s" % (image.shape, image_old.shape))
        self._image = image
        if update_pixels:
            self.update_pixels()

    def set_weight(self, weight, update_pixels=True):
        image = self.image
        if weight is not None:
            weight = np.asarray(weight, dtype='f8')
            assert len(weight.shape) == 2, "weight must be 2d"
            assert weight.shape == image.shape, \
                "image and weight must be same shape"
        else:
            weight = np.zeros(image.shape) + 1.0
        self._weight = weight
        if update_pixels:
            self.update_pixels()

    def set_jacobian(self, jacobian, update_pixels=True):
        if jacobian is None:
            cen = (np.array(self.image.shape) - 1.0) / 2.0
            jac = UnitJacobian(row=cen[0], col=cen[1])
        else:
            assert isinstance(jacobian, Jacobian), (
                "jacobian must be of type Jacobian, got %s" % type(jacobian))
            jac = jacobian.copy()
        self._jacobian = jac
        if update_pixels:
            self.update_pixels()

    def get_jacobian(self):
        j = self._jacobian.copy()
        j._data = self._get_view(self._jacobian._data)
        return j

    def set_psf(self, psf):
        if self.has_psf():
            del self._psf
        if psf is not None:
            assert isinstance(psf, Observation), \
                "psf must be of Observation, got %s" % type(psf)
            self._psf = psf

    def get_psf(self):
        if not self.has_psf():
            raise RuntimeError("this obs has no psf set")
        return self._psf

    def has_psf(self):
        return hasattr(self, '_psf')

    def get_psf_gmix(self):
        if not self.has_psf_gmix():
            raise RuntimeError("this obs has not psf set with a gmix")
        return self.psf.get_gmix()

    def has_psf_gmix(self):
        return self.has_psf() and self.psf.has_gmix()

    def set_gmix(self, gmix):
        if self.has_gmix():
            del self._gmix
        if gmix is not None:
            assert isinstance(gmix, GMix), \
                "gmix must be of type GMix, got %s" % type(gmix)
            self._gmix = gmix.copy()

    def get_gmix(self):
        if not self.has_gmix():
            raise RuntimeError("this obs has not gmix set")
        return self._gmix.copy()

    def has_gmix(self):
        return hasattr(self, '_gmix')

    def get_s2n(self):
        Isum, Vsum, Npix = self.get_s2n_sums()
        return Isum / np.sqrt(Vsum) if Vsum > 0.0 else -9999.0

    def get_s2n_sums(self):
        w = np.where(self.weight > 0)
        if w[0].size > 0:
            return (self.image[w].sum(), (1.0 / self.weight[w]).sum(),
                    w[0].size)
        return 0.0, 0.0, 0

    def copy(self, memo=None):
        def cp(name):
            return getattr(self, name).copy() \
                if getattr(self, 'has_' + name)() else None
        return Observation(
            self.image.copy(), weight=self.weight.copy(), bmask=cp('bmask'),
            ormask=cp('ormask'), noise=cp('noise'),
            gmix=self.gmix if self.has_gmix() else None,
            jacobian=self.jacobian,
            meta=copy.deepcopy(self._meta, memo=memo),
            psf=self.psf.copy() if self.has_psf() else None,
            mfrac=cp('mfrac'), store_pixels=self._store_pixels,
            ignore_zero_weight=self._ignore_zero_weight)

    def __copy__(self):
        return self.copy()

    def __deepcopy__(self, memo):
        result = self.copy(memo=memo)
        memo[id(self)] = result
        return result

    @property
    def store_pixels(self):
        return self._store_pixels

    @store_pixels.setter
    def store_pixels(self, store_pixels):
        changed = store_pixels != self._store_pixels
        self._store_pixels = store_pixels
        if changed:
            self.update_pixels()

    @property
    def ignore_zero_weight(self):
        return self._ignore_zero_weight

    @ignore_zero_weight.setter
    def ignore_zero_weight(self, ignore_zero_weight):
        changed = ignore_zero_weight != self._ignore_zero_weight
        self._ignore_zero_weight = ignore_zero_weight
        if changed:
            self.update_pixels()

    def update_pixels(self):
        if not self._store_pixels:
            self._pixels = None
            return
        pixels = make_pixels(self.image, self.weight, self._jacobian,
                             ignore_zero_weight=self._ignore_zero_weight)
        pixels.flags['WRITEABLE'] = False
        self._pixels = pixels

    def _get_view(self, data):
        view = data.view()
        view.flags['WRITEABLE'] = self._writeable
        return view

    def writeable(self):
        return self

    def __enter__(self):
        self._writeable = True
        return self

    def __exit__(self, exception_type, exception_value, traceback):
        self._writeable = False
        self.update_pixels()


class ObsList(list, MetadataMixin):
    """a list of Observations (the epochs of one band)"""

    def __init__(self, meta=None):
        super().__init__()
        self.set_meta(meta)

    def append(self, obs):
        assert isinstance(obs, Observation), \
            "obs should be of type Observation, got %s" % type(obs)
        super().append(obs)

    def get_s2n_sums(self):
        Isum = Vsum = 0.0
        Npix = 0
        for obs in self:
            a, b, c = obs.get_s2n_sums()
            Isum += a
            Vsum += b
            Npix += c
        return Isum, Vsum, Npix

    def get_s2n(self):
        Isum, Vsum, Npix = self.get_s2n_sums()
        return Isum / np.sqrt(Vsum) if Vsum > 0.0 else -9999.0

    def copy(self, memo=None):
        new = ObsList(meta=copy.deepcopy(self._meta, memo))
        for obs in self:
            new.append(obs.copy(memo=memo))
        return new

    def __setitem__(self, index, obs):
        assert isinstance(obs, Observation), \
            "obs should be of type Observation"
        super().__setitem__(index, obs)


class MultiBandObsList(list, MetadataMixin):
    """a list of ObsList, one per band"""

    def __init__(self, meta=None):
        super().__init__()
        self.set_meta(meta)

    def append(self, obs_list):
        assert isinstance(obs_list, ObsList), \
            'obs_list should be of type ObsList'
        super().append(obs_list)

    def get_s2n_sums(self):
        Isum = Vsum = 0.0
        Npix = 0
        for obslist in self:
            a, b, c = obslist.get_s2n_sums()
            Isum += a
            Vsum += b
            Npix += c
        return Isum, Vsum, Npix

    def get_s2n(self):
        Isum, Vsum, Npix = self.get_s2n_sums()
        return Isum / np.sqrt(Vsum) if Vsum > 0.0 else -9999.0

    def copy(self, memo=None):
        new = MultiBandObsList(meta=copy.deepcopy(self._meta, memo=memo))
        for obslist in self:
            new.append(obslist.copy(memo=memo))
        return new

    def __setitem__(self, index, obs_list):
        assert isinstance(obs_list, ObsList), \
            'obs_list should be of type ObsList'
        super().__setitem__(index, obs_list)


def get_mb_obs(obs_in):
    """convert an Observation or ObsList to a MultiBandObsList"""
    if isinstance(obs_in, Observation):
        obs_list = ObsList()
        obs_list.append(obs_in)
        obs = MultiBandObsList()
        obs.append(obs_list)
    elif isinstance(obs_in, ObsList):
        obs = MultiBandObsList()
        obs.append(obs_in)
    elif isinstance(obs_in, MultiBandObsList):
        obs = obs_in
    else:
        raise ValueError(
            'obs should be Observation, ObsList, or MultiBandObsList')
    return obs
