"""ORACLE-side synthetic input generators (test infrastructure only): seeded stick-figure heat/PAF maps in the
reference's 19-heat / 38-PAF COCO layout, and noise maps.  These are INPUTS for parity
tests, not a restatement of any reference function."""
import numpy as np

LIMB_PARTS = [(1, 2), (1, 5), (2, 3), (3, 4), (5, 6), (6, 7), (1, 8), (8, 9), (9, 10), (1, 11), (11, 12), (12, 13),
              (1, 0), (0, 14), (14, 16), (0, 15), (15, 17), (2, 16), (5, 17)]
LIMB_PAF_CH = [(12, 13), (20, 21), (14, 15), (16, 17), (22, 23), (24, 25), (0, 1), (2, 3), (4, 5), (6, 7), (8, 9),
               (10, 11), (28, 29), (30, 31), (34, 35), (32, 33), (36, 37), (18, 19), (26, 27)]
# canonical skeleton in a unit box (x, y), 18 COCO parts
_SKEL = np.array([[.50, .08], [.50, .22], [.36, .22], [.30, .40], [.27, .56], [.64, .22], [.70, .40], [.73, .56],
                  [.42, .55], [.41, .75], [.40, .95], [.58, .55], [.59, .75], [.60, .95], [.46, .05], [.54, .05],
                  [.41, .07], [.59, .07]])


def stick_figures(num_persons, seed, h=46, w=46, stride=8, sigma=7.0, paf_width=6.0, drop_prob=0.1):
    """Returns (heat [h,w,19] float32, paf [h,w,38] float32, keypoints [P,18,3] (x,y,visible) in image pixels)."""
    rs = np.random.RandomState(seed)
    H, W = h * stride, w * stride
    heat = np.zeros((h, w, 19), np.float32)
    paf = np.zeros((h, w, 38), np.float32)
    cnt = np.zeros((h, w, 19), np.float32)
    ys, xs = np.mgrid[0:h, 0:w]
    cx = xs * stride + stride / 2.0 - 0.5
    cy = ys * stride + stride / 2.0 - 0.5
    kps = np.zeros((num_persons, 18, 3))
    for p in range(num_persons):
        size = rs.uniform(0.28, 0.6) * min(H, W)
        ox = rs.uniform(0, W - size * 0.5)
        oy = rs.uniform(0, H - size)
        ang = rs.uniform(-0.35, 0.35)
        rot = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
        pts = (_SKEL - 0.5) @ rot.T * size + [ox + size * 0.5, oy + size * 0.5] + rs.normal(0, size * 0.015, (18, 2))
        vis = (rs.rand(18) > drop_prob) & (pts[:, 0] >= 0) & (pts[:, 0] < W) & (pts[:, 1] >= 0) & (pts[:, 1] < H)
        kps[p, :, :2] = pts
        kps[p, :, 2] = vis
        for j in range(18):
            if not vis[j]:
                continue
            g = np.exp(-((cx - pts[j, 0]) ** 2 + (cy - pts[j, 1]) ** 2) / (2 * sigma * sigma))
            heat[:, :, j] = np.maximum(heat[:, :, j], g * rs.uniform(0.75, 1.0))
        for l, (a, b) in enumerate(LIMB_PARTS):
            if not (vis[a] and vis[b]):
                continue
            v = pts[b] - pts[a]
            n = np.linalg.norm(v)
            if n < 1e-3:
                continue
            u = v / n
            rx, ry = cx - pts[a, 0], cy - pts[a, 1]
            along = rx * u[0] + ry * u[1]
            perp = np.abs(rx * u[1] - ry * u[0])
            m = (along >= 0) & (along <= n) & (perp <= paf_width)
            paf[:, :, LIMB_PAF_CH[l][0]][m] += u[0]
            paf[:, :, LIMB_PAF_CH[l][1]][m] += u[1]
            cnt[:, :, l][m] += 1
    for l in range(19):
        m = cnt[:, :, l] > 1
        for ch in LIMB_PAF_CH[l]:
            paf[:, :, ch][m] /= cnt[:, :, l][m]
    heat[:, :, 18] = 1.0 - heat[:, :, :18].max(axis=2)
    heat += rs.normal(0, 0.004, heat.shape).astype(np.float32)   # breaks exact ties / plateaus
    paf += rs.normal(0, 0.004, paf.shape).astype(np.float32)
    return heat.astype(np.float32), paf.astype(np.float32), kps


def noise_maps(seed, h=46, w=46, heat_scale=0.45, paf_scale=0.6):
    """Maps with the statistics of a random-weight network (thousands of peaks): the stress case."""
    rs = np.random.RandomState(seed)
    heat = (rs.standard_normal((h, w, 19)) * heat_scale).astype(np.float32)
    paf = (rs.standard_normal((h, w, 38)) * paf_scale).astype(np.float32)
    return heat, paf
