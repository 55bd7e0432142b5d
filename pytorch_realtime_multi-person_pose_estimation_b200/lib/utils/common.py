"""Result containers of the inference path, API-compatible with /root/reference/lib/utils/common.py
(CocoPart :5-24, Human :27-225, BodyPart :253-274, draw_humans :227-251, CocoColors/CocoPairs :276-284)."""
from enum import Enum

import cv2
import numpy as np


class CocoPart(Enum):
    Nose = 0
    Neck = 1
    RShoulder = 2
    RElbow = 3
    RWrist = 4
    LShoulder = 5
    LElbow = 6
    LWrist = 7
    RHip = 8
    RKnee = 9
    RAnkle = 10
    LHip = 11
    LKnee = 12
    LAnkle = 13
    REye = 14
    LEye = 15
    REar = 16
    LEar = 17
    Background = 18


class BodyPart:
    """One detected keypoint; x, y are normalised to [0, 1] by the network input size."""
    __slots__ = ('uidx', 'part_idx', 'x', 'y', 'score')

    def __init__(self, uidx, part_idx, x, y, score):
        self.uidx, self.part_idx, self.x, self.y, self.score = uidx, part_idx, x, y, score

    def get_part_name(self):
        return CocoPart(self.part_idx)

    def __repr__(self):
        return 'BodyPart:%d-(%.2f, %.2f) score=%.2f' % (self.part_idx, self.x, self.y, self.score)

    __str__ = __repr__


class Human:
    """A person: `body_parts` maps part index -> BodyPart; `score` is the mean limb/peak score."""
    __slots__ = ('body_parts', 'pairs', 'uidx_list', 'score')

    def __init__(self, pairs):
        self.pairs, self.uidx_list, self.body_parts, self.score = [], set(), {}, 0.0
        for pair in pairs:
            self.add_pair(pair)

    @staticmethod
    def _get_uidx(part_idx, idx):
        return '%d-%d' % (part_idx, idx)

    def add_pair(self, pair):
        self.pairs.append(pair)
        for pidx, idx, coord in ((pair.part_idx1, pair.idx1, pair.coord1), (pair.part_idx2, pair.idx2, pair.coord2)):
            uid = Human._get_uidx(pidx, idx)
            self.body_parts[pidx] = BodyPart(uid, pidx, coord[0], coord[1], pair.score)
            self.uidx_list.add(uid)

    def is_connected(self, other):
        return bool(self.uidx_list & other.uidx_list)

    def merge(self, other):
        for pair in other.pairs:
            self.add_pair(pair)

    def part_count(self):
        return len(self.body_parts)

    def get_max_score(self):
        return max(p.score for p in self.body_parts.values())

    def __repr__(self):
        return ' '.join(str(p) for p in self.body_parts.values())

    __str__ = __repr__


CocoColors = [[255, 0, 0], [255, 85, 0], [255, 170, 0], [255, 255, 0], [170, 255, 0], [85, 255, 0], [0, 255, 0],
              [0, 255, 85], [0, 255, 170], [0, 255, 255], [0, 170, 255], [0, 85, 255], [0, 0, 255], [85, 0, 255],
              [170, 0, 255], [255, 0, 255], [255, 0, 170], [255, 0, 85]]
CocoPairs = [(1, 2), (1, 5), (2, 3), (3, 4), (5, 6), (6, 7), (1, 8), (8, 9), (9, 10), (1, 11), (11, 12), (12, 13),
             (1, 0), (0, 14), (14, 16), (0, 15), (15, 17), (2, 16), (5, 17)]
CocoPairsRender = CocoPairs[:-2]


def draw_humans(npimg, humans, imgcopy=False):
    if imgcopy:
        npimg = np.copy(npimg)
    h, w = npimg.shape[:2]
    for human in humans:
        centers = {}
        for i, part in human.body_parts.items():
            if i >= CocoPart.Background.value:
                continue
            centers[i] = (int(part.x * w + 0.5), int(part.y * h + 0.5))
            cv2.circle(npimg, centers[i], 3, CocoColors[i], thickness=3, lineType=8, shift=0)
        for order, (a, b) in enumerate(CocoPairsRender):
            if a in centers and b in centers:
                cv2.line(npimg, centers[a], centers[b], CocoColors[order], 3)
    return npimg
