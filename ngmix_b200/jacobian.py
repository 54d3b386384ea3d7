"""
Jacobian / DiagonalJacobian / UnitJacobian: the host-side mirror of
ngmix/jacobian/jacobian.py (same constructor keywords, properties and record
layout).  The record `_data` is the 64-byte _jacobian_dtype the C ABI takes
as ngx_jacobian; v,u evaluation for whole images happens on the device
(ngx_make_pixels / ngx_make_coords and inside the fitting kernels).
"""
import numpy as np

from ._lib import JACOBIAN_DTYPE

__all__ = ['Jacobian', 'DiagonalJacobian', 'UnitJacobian']

_jacobian_dtype = JACOBIAN_DTYPE


class Jacobian(object):
    """
    The jacobian of the (row, col) -> (v, u) transformation about a centre.

    Send either row=, col=, dvdrow=, dvdcol=, dudrow=, dudcol= or
    x=, y=, dudx=, dudy=, dvdx=, dvdy= (or wcs= with those attributes).
    """

    def __init__(self, **kw):
        self._data = np.zeros(1, dtype=_jacobian_dtype)
        if 'x' in kw:
            if 'wcs' in kw:
                w = kw['wcs']
                d = (w.dvdy, w.dvdx, w.dudy, w.dudx)
            else:
                for k in ('x', 'y', 'dudx', 'dudy', 'dvdx', 'dvdy'):
                    if k not in kw:
                        raise ValueError("missing keyword: '%s'" % k)
                d = (kw['dvdy'], kw['dvdx'], kw['dudy'], kw['dudx'])
            self._finish_init(kw['y'], kw['x'], *d)
        elif 'row' in kw:
            if 'wcs' in kw:
                w = kw['wcs']
                d = (w.dvdy, w.dvdx, w.dudy, w.dudx)
            else:
                for k in ('row', 'col', 'dvdrow', 'dvdcol', 'dudrow',
                          'dudcol'):
                    if k not in kw:
                        raise ValueError("missing keyword: '%s'" % k)
                d = (kw['dvdrow'], kw['dvdcol'], kw['dudrow'], kw['dudcol'])
            self._finish_init(kw['row'], kw['col'], *d)
        else:
            raise ValueError("send by row,col or x,y")

    def _finish_init(self, row0, col0, dvdrow, dvdcol, dudrow, dudcol):
        d = self._data
        d['row0'] = row0
        d['col0'] = col0
        d['dvdrow'] = dvdrow
        d['dvdcol'] = dvdcol
        d['dudrow'] = dudrow
        d['dudcol'] = dudcol
        d['det'] = dvdrow * dudcol - dvdcol * dudrow
        d['scale'] = np.sqrt(np.abs(d['det']))

    def get_data(self):
        return self._data

    def get_cen(self):
        return self._data['row0'][0], self._data['col0'][0]

    def get_row0(self):
        return self._data['row0'][0]

    def get_col0(self):
        return self._data['col0'][0]

    def get_dvdrow(self):
        return self._data['dvdrow'][0]

    def get_dvdcol(self):
        return self._data['dvdcol'][0]

    def get_dudrow(self):
        return self._data['dudrow'][0]

    def get_dudcol(self):
        return self._data['dudcol'][0]

    def get_det(self):
        return self._data['det'][0]

    def get_scale(self):
        return self._data['scale'][0]

    def get_area(self):
        return self.scale ** 2

    def get_vu(self, row, col):
        """v,u for scalar or array row,col (jacobian_nb.py:4-16)"""
        d = self._data
        rowdiff = row - d['row0'][0]
        coldiff = col - d['col0'][0]
        v = d['dvdrow'][0] * rowdiff + d['dvdcol'][0] * coldiff
        u = d['dudrow'][0] * rowdiff + d['dudcol'][0] * coldiff
        return v, u

    def get_rowcol(self, v, u):
        """row,col for scalar or array v,u (jacobian_nb.py:19-31)"""
        d = self._data
        rowdiff = d['dudcol'][0] * v - d['dvdcol'][0] * u
        coldiff = -d['dudrow'][0] * v + d['dvdrow'][0] * u
        row = d['row0'][0] + rowdiff / d['det'][0]
        col = d['col0'][0] + coldiff / d['det'][0]
        return row, col

    def __call__(self, row, col):
        return self.get_vu(row, col)

    cen = property(fget=get_cen)
    row0 = property(fget=get_row0)
    col0 = property(fget=get_col0)
    dvdrow = property(fget=get_dvdrow)
    dvdcol = property(fget=get_dvdcol)
    dudrow = property(fget=get_dudrow)
    dudcol = property(fget=get_dudcol)
    det = property(fget=get_det)
    scale = property(fget=get_scale)
    area = property(fget=get_area)

    def set_cen(self, **kw):
        if 'row' in kw:
            self._data['row0'] = kw['row']
            self._data['col0'] = kw['col']
        elif 'x' in kw:
            self._data['row0'] = kw['y']
            self._data['col0'] = kw['x']
        else:
            raise ValueError("expected row=,col= or x=,y=")

    def copy(self):
        return Jacobian(row=self.row0, col=self.col0, dudrow=self.dudrow,
                        dudcol=self.dudcol, dvdrow=self.dvdrow,
                        dvdcol=self.dvdcol)

    def __copy__(self):
        return self.copy()

    def __deepcopy__(self, memo):
        result = self.copy()
        memo[id(self)] = result
        return result

    def __eq__(self, jacobian):
        if not isinstance(jacobian, Jacobian):
            raise ValueError(f'expected Jacobian, got {type(jacobian)}')
        return bool(np.all(self.get_data() == jacobian.get_data()))

    def __repr__(self):
        fmt = ('ngmix.Jacobian(row=%r, col=%r, dvdrow=%r, dvdcol=%r, '
               'dudrow=%r, dudcol=%r)')
        return fmt % (self.row0, self.col0, self.dvdrow, self.dvdcol,
                      self.dudrow, self.dudcol)


class DiagonalJacobian(Jacobian):
    """u varies with column/x and v with row/y, by `scale`"""

    def __init__(self, scale=1.0, **kw):
        if 'x' in kw:
            assert 'y' in kw, "send both x= and y="
            super().__init__(x=kw['x'], y=kw['y'], dudx=scale, dudy=0.0,
                             dvdx=0.0, dvdy=scale)
        elif 'row' in kw:
            assert 'col' in kw, "send both row= and col="
            super().__init__(row=kw['row'], col=kw['col'], dvdrow=scale,
                             dvdcol=0.0, dudrow=0.0, dudcol=scale)
        else:
            raise ValueError("expected row=,col= or x=,y=")


class UnitJacobian(DiagonalJacobian):
    """diagonal jacobian with scale 1"""

    def __init__(self, **kw):
        super().__init__(scale=1.0, **kw)
