"""ctypes binding of the CPU oracle (oracle/_build/libvpp_oracle.so) + host image helper.

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIBS = {}


class VoImg(C.Structure):
    _fields_ = [("base", C.c_void_p), ("nrows", C.c_int32), ("ncols", C.c_int32), ("pitch", C.c_int32),
                ("border", C.c_int32), ("elem", C.c_int32)]


class VoLkParams(C.Structure):
    _fields_ = [("nlevels", C.c_int32), ("min_scale", C.c_int32), ("winsize", C.c_int32), ("max_iter", C.c_int32),
                ("grad_is_float", C.c_int32), ("err_mode", C.c_int32), ("gate_on_max_err", C.c_int32),
                ("min_ev", C.c_float), ("delta", C.c_float), ("max_err", C.c_float), ("factor", C.c_float),
                ("pred_div", C.c_float)]


def load(omp=False):
    key = "omp" if omp else "serial"
    if key not in _LIBS:
        path = os.path.join(ROOT, "oracle", "_build", "libvpp_oracle_omp.so" if omp else "libvpp_oracle.so")
        lib = C.CDLL(path)
        P = C.POINTER
        I = P(VoImg)
        lib.vo_layout.argtypes = [C.c_int] * 5 + [P(C.c_int), P(C.c_int64), P(C.c_int64)]
        lib.vo_pw_add_i32.argtypes = [I, I, I]
        lib.vo_fill.argtypes = [I, C.c_void_p, C.c_int]
        lib.vo_copy.argtypes = [I, I, C.c_int]
        lib.vo_fill_border_value.argtypes = [I, C.c_void_p]
        lib.vo_fill_border_mirror.argtypes = [I]
        lib.vo_fill_border_closest.argtypes = [I]
        lib.vo_sum_i32.argtypes = [I, C.c_int]
        lib.vo_sum_i32.restype = C.c_int64
        lib.vo_box5x5_u8.argtypes = [I, I, C.c_int]
        lib.vo_box5x5_i32.argtypes = [I, I]
        lib.vo_scharr_u8.argtypes = [I, I, C.c_int]
        lib.vo_rgb_to_graylevel.argtypes = [I, I]
        lib.vo_lowpass_sub2.argtypes = [I, I, C.c_int]
        lib.vo_lowpass.argtypes = [I, I, C.c_int]
        lib.vo_fast9_u8.argtypes = [I, C.c_int, I, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        lib.vo_fast9_score.argtypes = [I, C.c_int, C.c_int, C.c_int]
        lib.vo_interp_u8.argtypes = [I, C.c_float, C.c_float]
        lib.vo_lk_match_u8.argtypes = [I, I, I, P(VoLkParams), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.vo_semi_dense_flow.argtypes = [I, I, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.vo_flow_error_stats.argtypes = [I, I, C.c_void_p, I]
        lib.vo_lbp_u8.argtypes = [I, I]
        lib.vo_local_maxima_filter.argtypes = [I]
        lib.vo_fast9_blockwise_rank.argtypes = [I, C.c_int, I, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        lib.vo_lk_match_oriented_u8.argtypes = [I, I, I, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p]
        lib.vo_num_threads.restype = C.c_int
        _LIBS[key] = lib
    return _LIBS[key]


PIXEL_TYPES = {"u8": (np.uint8, 1), "i8": (np.int8, 1), "vuchar3": (np.uint8, 3), "vuchar4": (np.uint8, 4), "i32": (np.int32, 1),
               "f32": (np.float32, 1), "vint2": (np.int32, 2), "vfloat2": (np.float32, 2), "vfloat3": (np.float32, 3)}


class HostImage:
    """image2d<V> in host memory with the reference layout (imageNd.hpp:151-196)."""

    def __init__(self, nrows, ncols, pixel="u8", border=0, aligned=128, data=None, fill_border=None, border_value=0):
        self.pixel = pixel
        self.dtype, self.channels = PIXEL_TYPES[pixel]
        self.elem = np.dtype(self.dtype).itemsize * self.channels
        self.nrows, self.ncols, self.border, self.aligned = nrows, ncols, border, aligned
        pitch, total, origin = C.c_int(), C.c_int64(), C.c_int64()
        assert load().vo_layout(nrows, ncols, self.elem, border, aligned, C.byref(pitch), C.byref(total), C.byref(origin)) == 0
        self.pitch, self.total, self.origin = pitch.value, total.value, origin.value
        raw = np.zeros(self.total + aligned, dtype=np.uint8)
        off = (-raw.ctypes.data) % aligned
        self.buf = raw[off:off + self.total]
        self._raw = raw
        self.desc = VoImg(self.buf.ctypes.data + self.origin, nrows, ncols, self.pitch, border, self.elem)
        if data is not None:
            self.set(data)
        if fill_border == "mirror":
            load().vo_fill_border_mirror(self.ptr())
        elif fill_border == "value":
            v = np.full(self.channels, border_value, dtype=self.dtype)
            load().vo_fill_border_value(self.ptr(), v.ctypes.data)

    def ptr(self):
        return C.byref(self.desc)

    def _frame_view(self, b):
        """numpy view of rows [-b, nrows+b) x cols [-b, ncols+b)."""
        start = self.origin - b * self.pitch - b * self.elem
        rows = self.nrows + 2 * b
        cols = self.ncols + 2 * b
        item = np.dtype(self.dtype).itemsize
        shape = (rows, cols, self.channels)
        strides = (self.pitch, self.elem, item)
        v = np.ndarray(shape, dtype=self.dtype, buffer=self.buf, offset=start, strides=strides)
        return v[:, :, 0] if self.channels == 1 else v

    def view(self, with_border=False):
        return self._frame_view(self.border if with_border else 0)

    def set(self, data, with_border=False):
        self.view(with_border)[...] = np.asarray(data, dtype=self.dtype).reshape(self.view(with_border).shape)
        return self

    def get(self, with_border=False):
        return np.array(self.view(with_border))


def desc_array(imgs):
    arr = (VoImg * len(imgs))()
    for i, im in enumerate(imgs):
        arr[i] = im.desc
    return arr
