"""ORACLE (test infrastructure only - never imported by the product path).

CPU restatement, in numpy float32, of the heat-map peak extraction of the reference:

  * find_peaks            /root/reference/lib/utils/paf_to_pose.py:25-38
  * compute_resized_coords /root/reference/lib/utils/paf_to_pose.py:41-64
  * NMS (refine branch)    /root/reference/lib/utils/paf_to_pose.py:67-145

Third-party arithmetic restated here (not vendored in /root/reference):
  * scipy.ndimage.maximum_filter(footprint=cross3, mode='reflect') (SciPy, unpinned in requirements.txt:6;
    container 1.18.1): a pixel is a peak iff it is >= its in-bounds 4-neighbours (reflect duplicates the edge
    pixel itself) and > threshold.
  * cv2.resize(patch, fx=fy=8, INTER_CUBIC) on float32 (OpenCV is not listed in requirements.txt; container
    4.13.0).  Restated from OpenCV's published algorithm (modules/imgproc/src/resize.cpp: interpolateCubic with
    A=-0.75, separable, horizontal pass `((s0*a0+s1*a1)+s2*a2)+s3*a3`, vertical pass
    `s0*b0+(s1*b1+(s2*b2+s3*b3))`, replicate border, no FMA).  Pinned (tests/test_oracle_nms.py): bit-exact
    against cv2 with IPP disabled; against the default cv2 build of this container (which routes float cubic
    resize through Intel IPP, proprietary arithmetic) max |diff| <= 5e-7 with identical arg-max on every fixture.

Only up-sampling factor 8 (cfg.MODEL.DOWNSAMPLE, /root/reference/lib/config/default.py:41) is supported.
"""
import numpy as np

F32 = np.float32
UPS = 8
WIN = 2  # paf_to_pose.py:100


def _cubic_coeffs(x):
    """OpenCV interpolateCubic, float32, A = -0.75."""
    A = F32(-0.75)
    x = F32(x)
    one = F32(1)
    c0 = ((A * (x + one) - F32(5) * A) * (x + one) + F32(8) * A) * (x + one) - F32(4) * A
    c1 = ((A + F32(2)) * x - (A + F32(3))) * x * x + one
    y = one - x
    c2 = ((A + F32(2)) * y - (A + F32(3))) * y * y + one
    c3 = one - c0 - c1 - c2
    return np.array([c0, c1, c2, c3], dtype=F32)


def cubic_tables():
    """(coeff[8,4] float32, first_tap_offset[8] int) for destination phase d = dst % 8:
    src = (dst + 0.5)/8 - 0.5 ; taps at floor(src) - 1 + {0..3}."""
    tab, off = [], []
    for d in range(UPS):
        fx = F32((d + 0.5) * (1.0 / UPS) - 0.5)
        sx = int(np.floor(fx))
        tab.append(_cubic_coeffs(F32(fx - F32(sx))))
        off.append(sx - 1)
    return np.array(tab, dtype=F32), np.array(off, dtype=np.int64)


_TAB, _OFF = cubic_tables()


def upsample8_cubic(patch):
    """cv2.resize(patch, None, fx=8, fy=8, interpolation=cv2.INTER_CUBIC) for a small float32 patch."""
    patch = np.ascontiguousarray(patch, dtype=F32)
    h, w = patch.shape
    X = np.arange(w * UPS)
    ix = np.clip((X // UPS + _OFF[X % UPS])[:, None] + np.arange(4)[None, :], 0, w - 1)  # [W,4]
    ax = _TAB[X % UPS]                                                                     # [W,4]
    hor = patch[:, ix[:, 0]] * ax[:, 0]
    for j in (1, 2, 3):
        hor = hor + patch[:, ix[:, j]] * ax[:, j]
    Y = np.arange(h * UPS)
    iy = np.clip((Y // UPS + _OFF[Y % UPS])[:, None] + np.arange(4)[None, :], 0, h - 1)  # [H,4]
    by = _TAB[Y % UPS]
    out = hor[iy[:, 3]] * by[:, 3:4]
    for j in (2, 1, 0):
        out = hor[iy[:, j]] * by[:, j:j + 1] + out
    return out.astype(F32)


def find_peaks(thresh, img):
    """paf_to_pose.py:25-38.  Returns [[x, y], ...] in raster (y-major) order."""
    img = np.asarray(img)
    h, w = img.shape
    ge = np.ones((h, w), dtype=bool)
    ge[1:, :] &= img[1:, :] >= img[:-1, :]
    ge[:-1, :] &= img[:-1, :] >= img[1:, :]
    ge[:, 1:] &= img[:, 1:] >= img[:, :-1]
    ge[:, :-1] &= img[:, :-1] >= img[:, 1:]
    binary = ge & (img > thresh)
    ys, xs = np.nonzero(binary)
    return np.stack([xs, ys], axis=1)


def nms(heatmaps, thresh, num_keypoints=18):
    """NMS(..., upsampFactor=8, bool_refine_center=True, bool_gaussian_filt=False), paf_to_pose.py:67-145.
    heatmaps: (h, w, >=num_keypoints) float32.  Returns a list of num_keypoints arrays (k, 4) float64
    (x, y, score, id)."""
    out = []
    cnt = 0
    h, w = heatmaps.shape[:2]
    for joint in range(num_keypoints):
        m = np.ascontiguousarray(heatmaps[:, :, joint], dtype=F32)
        coords = find_peaks(thresh, m)
        peaks = np.zeros((len(coords), 4))
        for i, (px, py) in enumerate(coords):
            x_min, y_min = max(0, px - WIN), max(0, py - WIN)
            x_max, y_max = min(w - 1, px + WIN), min(h - 1, py + WIN)
            up = upsample8_cubic(m[y_min:y_max + 1, x_min:x_max + 1])
            loc = np.unravel_index(up.argmax(), up.shape)  # first maximum, row-major
            # (p+0.5)*8-0.5 + (argmax - ((p-min)+0.5)*8+0.5)  ==  8*min + argmax   (paf_to_pose.py:126-139)
            cx = (px + 0.5) * UPS - 0.5 + (loc[1] - ((px - x_min + 0.5) * UPS - 0.5))
            cy = (py + 0.5) * UPS - 0.5 + (loc[0] - ((py - y_min + 0.5) * UPS - 0.5))
            peaks[i] = (cx, cy, up[loc], cnt)
            cnt += 1
        out.append(peaks)
    return out


def joint_list_from_nms(per_joint):
    """paf_to_pose.py:376-378: float32 [P, 5] rows (x, y, score, id, part)."""
    rows = [tuple(p) + (j,) for j, peaks in enumerate(per_joint) for p in peaks]
    return np.array(rows, dtype=np.float32).reshape(-1, 5)
