"""Oracle backend of vpp_b200.video_extruder's `ops` interface (test infrastructure)."""
import numpy as np

from . import oracle as orc


class OracleOps:
    def __init__(self, lib=None):
        self.o = lib or orc.load()

    def _img(self, frame, border=0):
        return orc.HostImage(frame.shape[0], frame.shape[1], "u8", border=border, data=frame, fill_border="mirror" if border else None)

    def flow(self, kps, f1, f2, winsize, nscales, propagation, patchsize):
        n = len(kps)
        k = np.ascontiguousarray(kps, np.int32)
        pos, dist, valid = np.zeros((n, 2), np.int32), np.zeros(n, np.int32), np.zeros(n, np.uint8)
        h1, h2 = self._img(f1), self._img(f2)
        self.o.vo_semi_dense_flow(h1.ptr(), h2.ptr(), k.ctypes.data, n, winsize, nscales, 0, propagation, patchsize, pos.ctypes.data, dist.ctypes.data,
                                  valid.ctypes.data)
        return pos, dist, valid.astype(bool)

    def fast_scores(self, frame, th, pts):
        h = self._img(frame, 3)
        return np.array([self.o.vo_fast9_score(h.ptr(), th, int(r), int(c)) for r, c in pts], np.int32)

    def fast_blockwise(self, frame, th, block_size, mask):
        h, hm = self._img(frame, 3), self._img(np.ascontiguousarray(mask))
        k = np.zeros((frame.size, 2), np.int32)
        n = self.o.vo_fast9_u8(h.ptr(), th, hm.ptr(), 2, block_size, 0, k.ctypes.data, None, len(k))
        return k[:n]
