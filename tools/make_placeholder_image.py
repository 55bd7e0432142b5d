#!/usr/bin/env python
"""Writes readme/ski.jpg: a SYNTHETIC stand-in with the dimensions of the picture the reference demo reads
(`./readme/ski.jpg`, demo/picture_demo.py:51 - 674 x 712 pixels).  The reference's photograph is not redistributed;
this file only has to exist, decode as a 3-channel image of that size, and exercise the non-square resize + padding
path (368/674 scale -> 368 x 389 -> padded to 368 x 392)."""
import os

import cv2
import numpy as np

H, W = 674, 712
rs = np.random.RandomState(51)
yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
img = np.stack([180 + 40 * yy / H, 190 + 30 * xx / W, 205 - 35 * yy / H], axis=2)          # sky / snow gradient
for _ in range(24):                                                                       # soft blobs
    cx, cy, r = rs.uniform(0, W), rs.uniform(0, H), rs.uniform(15, 120)
    img += np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * r * r))[:, :, None] * rs.uniform(-90, 90, 3)
img = np.clip(img, 0, 255).astype(np.uint8)
for _ in range(3):                                                                        # a few schematic figures
    ox, oy, s = rs.uniform(80, W - 200), rs.uniform(60, H - 360), rs.uniform(180, 300)
    pt = lambda x, y: (int(ox + x * s), int(oy + y * s))
    col = tuple(int(v) for v in rs.randint(20, 120, 3))
    cv2.circle(img, pt(.5, .08), int(.07 * s), col, -1)
    for a, b in (((.5, .15), (.5, .55)), ((.5, .22), (.3, .45)), ((.5, .22), (.7, .45)), ((.5, .55), (.4, .95)),
                 ((.5, .55), (.6, .95))):
        cv2.line(img, pt(*a), pt(*b), col, int(.05 * s))
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "readme", "ski.jpg")
cv2.imwrite(out, img, [cv2.IMWRITE_JPEG_QUALITY, 90])
print(out, img.shape)
