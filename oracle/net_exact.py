"""ORACLE (test infrastructure only).  The rtpose VGG19 forward pass (/root/reference/lib/network/rtpose_vgg.py:158-198)
in fp32 with a DEFINED accumulation order (oracle/conv_exact.c): the CPU side of the bit-exact end-to-end parity tests.

`net_port.forward` (torch CPU ops = the reference's own arithmetic, whatever order oneDNN picks on this CPU) stays the
restatement that is pinned to the reference; this module is pinned to it in turn (tests/test_oracle.py, < 1e-4 on the
reference's golden 368x368 output) and exists because "identical keypoint assignments" can only be asserted between
two pipelines whose maps are identical: NMS / PAF scoring on random-weight maps flips decisions under 3e-5
perturbations (measured: 0 of 18 trials kept the same persons), so tolerance-level agreement of the maps is not enough.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import net_port

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
_FP = ctypes.POINTER(ctypes.c_float)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libconv_exact.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", _HERE, "libconv_exact.so"], check=True)
        L = ctypes.CDLL(path)
        L.exact_conv.argtypes = [_FP] + [ctypes.c_int] * 6 + [_FP, _FP, ctypes.c_int, ctypes.c_int, _FP] + [ctypes.c_int] * 3
        L.exact_conv.restype = ctypes.c_int
        L.exact_maxpool.argtypes = [_FP] + [ctypes.c_int] * 4 + [_FP]
        L.exact_maxpool.restype = None
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_FP)


def conv(x, w, b, relu, out=None, out_off=0, in_off=0, cin=None):
    """x: NHWC float32 (channel slice [in_off, in_off+cin)); w: OIHW; returns / fills NHWC `out` at out_off."""
    n, h, wd, cs = x.shape
    cout, wcin, ks, _ = w.shape
    cin = wcin if cin is None else cin
    assert cin == wcin and x.flags.c_contiguous and x.dtype == np.float32
    w = np.ascontiguousarray(w, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    if out is None:
        out = np.empty((n, h, wd, cout), np.float32)
    rc = lib().exact_conv(_p(x), n, h, wd, cs, in_off, cin, _p(w), _p(b), cout, ks, _p(out), out.shape[3], out_off,
                          int(relu))
    assert rc == 0
    return out


def maxpool(x):
    n, h, w, c = x.shape
    out = np.empty((n, h // 2, w // 2, c), np.float32)
    lib().exact_maxpool(_p(x), n, h, w, c, _p(out))
    return out


def forward(sd, x_nchw):
    """x: float32 [N,3,H,W] (numpy or torch) -> ((paf, heat), saved_for_loss[12]) as NCHW numpy arrays."""
    x = np.ascontiguousarray(np.asarray(x_nchw, dtype=np.float32).transpose(0, 2, 3, 1))
    g = lambda k: sd[k].numpy() if hasattr(sd[k], "numpy") else np.asarray(sd[k])
    idx = 0
    for item in net_port.TRUNK:
        if item == "P":
            x = maxpool(x)
            idx += 1
            continue
        x = conv(x, g("model0.%d.weight" % idx), g("model0.%d.bias" % idx), True)
        idx += 2
    n, h, w, _ = x.shape
    cat = np.zeros((n, h, w, 38 + 19 + 128), np.float32)      # torch.cat([paf, heat, feat], 1), rtpose_vgg.py:165
    cat[:, :, :, 57:] = x
    saved = []
    for stage in range(1, net_port.NUM_STAGES + 1):
        outs = []
        for branch in (1, 2):
            layers = net_port.stage_layers(stage, branch)
            y = None
            for li, (_, _, k) in enumerate(layers):
                wk, bk = ("model%d_%d.%d.weight" % (stage, branch, 2 * li)), ("model%d_%d.%d.bias" % (stage, branch, 2 * li))
                last = li == len(layers) - 1
                if li == 0:
                    y = conv(cat, g(wk), g(bk), not last, in_off=57 if stage == 1 else 0, cin=128 if stage == 1 else 185)
                else:
                    y = conv(y, g(wk), g(bk), not last)
            outs.append(y)
        cat[:, :, :, 0:38] = outs[0]
        cat[:, :, :, 38:57] = outs[1]
        saved += [np.ascontiguousarray(o.transpose(0, 3, 1, 2)) for o in outs]
    return (saved[-2], saved[-1]), saved
