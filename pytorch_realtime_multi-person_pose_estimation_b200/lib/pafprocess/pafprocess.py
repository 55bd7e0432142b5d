"""Drop-in for the SWIG module built from /root/reference/lib/pafprocess/pafprocess.i: same function names and
argument meaning (pafprocess.h:53-59), backed by the CUDA kernels in libb200pose.so.

process_paf(peaks[1,P,5] f32, heat_up[H,W,19] f32, paf_up[H,W,38] f32) -> 0
getters read the state left by the last process_paf call (global, like the reference's pafprocess.cpp:12-13)."""
import numpy as np

from ... import _native as nat


def _as_f32_3d(a, name):
    # numpy.i IN_ARRAY3 typemap semantics: float32, 3-D, made contiguous
    if not isinstance(a, np.ndarray) or a.dtype != np.float32:
        raise TypeError("%s: array of type 'float' (float32) required" % name)
    if a.ndim != 3:
        raise ValueError("%s: array must have 3 dimensions, given array has %d" % (name, a.ndim))
    return np.ascontiguousarray(a)


def process_paf(peaks, heatmap, pafmap):
    peaks, heatmap, pafmap = _as_f32_3d(peaks, "peaks"), _as_f32_3d(heatmap, "heatmap"), _as_f32_3d(pafmap, "pafmap")
    rc = nat.lib().process_paf(*peaks.shape, peaks.ctypes.data, *heatmap.shape, heatmap.ctypes.data, *pafmap.shape,
                               pafmap.ctypes.data)
    nat.check(rc, "process_paf")
    return 0


def get_num_humans():
    return nat.lib().get_num_humans()


def get_part_cid(human_id, part_id):
    return nat.lib().get_part_cid(human_id, part_id)


def get_score(human_id):
    return nat.lib().get_score(human_id)


def get_part_x(cid):
    return nat.lib().get_part_x(cid)


def get_part_y(cid):
    return nat.lib().get_part_y(cid)


def get_part_score(cid):
    return nat.lib().get_part_score(cid)
