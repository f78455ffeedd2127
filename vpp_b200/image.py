"""image2d<V> on pitched HBM — the Python host mirror of vpp/core/imageNd.hh:42-168.

Only the container lives here; every pixel operation goes through the C-ABI (capi.lib).
"""
import ctypes as C

import numpy as np

from . import capi

DEFAULT_ALIGNMENT = 128  # reference default is 16/32 (imageNd.hpp:10-18); B200 rows are 128-B aligned

# pixel type name -> (numpy scalar dtype, channels)
PIXEL_TYPES = {
    "u8": (np.uint8, 1),       # image2d<unsigned char>
    "i8": (np.int8, 1),        # image2d<char>
    "vuchar3": (np.uint8, 3),  # image2d<vuchar3>
    "vuchar4": (np.uint8, 4),  # image2d<vuchar4> (RGBA / BGRA surfaces)
    "i32": (np.int32, 1),      # image2d<int>
    "f32": (np.float32, 1),    # image2d<float>
    "vint2": (np.int32, 2),    # image2d<vint2>
    "vfloat2": (np.float32, 2),  # image2d<vfloat2>
}


def layout(nrows, ncols, elem_bytes, border=0, aligned=DEFAULT_ALIGNMENT):
    """(pitch, total_bytes, origin_offset) per imageNd.hpp:151-196 (host arithmetic, no GPU)."""
    pitch, total, origin = C.c_int32(), C.c_int64(), C.c_int64()
    capi.check(capi.lib.vppb_layout(nrows, ncols, elem_bytes, border, aligned, C.byref(pitch), C.byref(total), C.byref(origin)))
    return pitch.value, total.value, origin.value


class Box2d:
    """boxNd<2> (boxNd.hh:11-74): inclusive integer box [p1, p2], coordinates (row, col)."""

    def __init__(self, p1, p2):
        self.p1 = (int(p1[0]), int(p1[1]))
        self.p2 = (int(p2[0]), int(p2[1]))

    @property
    def nrows(self):
        return self.p2[0] - self.p1[0] + 1

    @property
    def ncols(self):
        return self.p2[1] - self.p1[1] + 1

    def has(self, p):
        return self.p1[0] <= p[0] <= self.p2[0] and self.p1[1] <= p[1] <= self.p2[1]

    def __eq__(self, o):
        return isinstance(o, Box2d) and self.p1 == o.p1 and self.p2 == o.p2

    def __repr__(self):
        return "Box2d(%s, %s)" % (self.p1, self.p2)


def make_box2d(nrows, ncols):
    return Box2d((0, 0), (nrows - 1, ncols - 1))


class Image2d:
    """image2d<V>(nrows, ncols, _border=, _aligned=) in device memory.

    Copies share the buffer (imageNd.hpp:77-87): `b = a.share()`; `a | box` (subimage) aliases pixels.
    """

    def __init__(self, nrows, ncols, pixel="u8", border=0, aligned=DEFAULT_ALIGNMENT, _desc=None, _owner=None):
        self.pixel = pixel
        self.dtype, self.channels = PIXEL_TYPES[pixel]
        self.elem_bytes = np.dtype(self.dtype).itemsize * self.channels
        if _desc is not None:
            self.desc = _desc
            self._owner = _owner
            return
        self.desc = capi.VppbImg()
        capi.check(capi.lib.vppb_alloc(C.byref(self.desc), nrows, ncols, self.elem_bytes, border, aligned))
        self._owner = _Owner(self.desc)

    # --- accessors (imageNd.hh:88-161)
    nrows = property(lambda s: s.desc.nrows)
    ncols = property(lambda s: s.desc.ncols)
    pitch = property(lambda s: s.desc.pitch)
    border = property(lambda s: s.desc.border)
    alignment = property(lambda s: s.desc.align)

    @property
    def domain(self):
        return make_box2d(self.nrows, self.ncols)

    @property
    def domain_with_border(self):
        b = self.border
        return Box2d((-b, -b), (self.nrows - 1 + b, self.ncols - 1 + b))

    def has(self, p):
        return self.domain.has(p)

    def ptr(self):
        return C.byref(self.desc)

    def share(self):
        return Image2d(0, 0, self.pixel, _desc=self.desc, _owner=self._owner)

    def subimage(self, box):
        out = capi.VppbImg()
        capi.check(capi.lib.vppb_subimage(self.ptr(), box.p1[0], box.p1[1], box.p2[0], box.p2[1], C.byref(out)))
        return Image2d(0, 0, self.pixel, _desc=out, _owner=self._owner)

    __or__ = subimage  # img | box (imageNd.hh:173-177)

    # --- host <-> device
    def _host_shape(self, with_border):
        b = self.border if with_border else 0
        shp = (self.nrows + 2 * b, self.ncols + 2 * b)
        return shp + ((self.channels,) if self.channels > 1 else ())

    def upload(self, host, with_border=False, stream=None):
        """host: array of the domain, or of the domain + border frame when with_border."""
        a = np.ascontiguousarray(host, dtype=self.dtype)
        assert a.shape == self._host_shape(with_border), (a.shape, self._host_shape(with_border))
        b = self.border if with_border else 0
        host_pitch = (self.ncols + 2 * b) * self.elem_bytes
        origin = a.ctypes.data + b * host_pitch + b * self.elem_bytes
        capi.check(capi.lib.vppb_upload(self.ptr(), origin, host_pitch, 1 if with_border else 0, stream))
        capi.check(capi.lib.vppb_sync(stream))
        return self

    def download(self, with_border=False, stream=None):
        a = np.empty(self._host_shape(with_border), dtype=self.dtype)
        b = self.border if with_border else 0
        host_pitch = (self.ncols + 2 * b) * self.elem_bytes
        origin = a.ctypes.data + b * host_pitch + b * self.elem_bytes
        capi.check(capi.lib.vppb_download(self.ptr(), origin, host_pitch, 1 if with_border else 0, stream))
        capi.check(capi.lib.vppb_sync(stream))
        return a

    @classmethod
    def from_host(cls, host, pixel, border=0, aligned=DEFAULT_ALIGNMENT):
        host = np.asarray(host)
        img = cls(host.shape[0], host.shape[1], pixel, border=border, aligned=aligned)
        return img.upload(host)


class _Owner:
    """shared_ptr<void> deleter of imageNd.hpp:177-180."""

    def __init__(self, desc):
        self._desc = capi.VppbImg()
        C.memmove(C.byref(self._desc), C.byref(desc), C.sizeof(capi.VppbImg))

    def __del__(self):
        try:
            if self._desc.alloc:
                capi.lib.vppb_free(C.byref(self._desc))
        except Exception:
            pass
