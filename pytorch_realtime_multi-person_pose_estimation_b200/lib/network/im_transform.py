"""Drop-in for the one function of /root/reference/lib/network/im_transform.py the inference path uses."""
import cv2
import numpy as np


def _factor_closest(num, factor, is_ceil=True):
    q = float(num) / factor
    return int(np.ceil(q) if is_ceil else np.floor(q)) * factor


def crop_with_factor(im, dest_size=None, factor=32, is_ceil=True):
    """im_transform.py:119-134: scale so the short side is dest_size (bilinear), zero-pad bottom/right to a
    multiple of `factor`.  Returns (padded image, scale, resized shape)."""
    scale = float(dest_size) / min(im.shape[0], im.shape[1])
    resized = cv2.resize(im, None, fx=scale, fy=scale)
    h, w, c = resized.shape
    out = np.zeros([_factor_closest(h, factor, is_ceil), _factor_closest(w, factor, is_ceil), c], dtype=resized.dtype)
    out[:h, :w, :] = resized
    return out, scale, resized.shape
