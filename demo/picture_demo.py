#!/usr/bin/env python
"""Single-picture demo on the B200 path (same flow and CLI as the reference's demo/picture_demo.py: load weights,
get_outputs, paf_to_pose_cpp, draw_humans, write result.png).  Without --weight a seeded random-init model is used
(no checkpoint ships offline); --image defaults to a synthetic frame.  Run from the repo root."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import cv2  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402

from evaluate.coco_eval import get_outputs  # noqa: E402
from lib.config import cfg, update_config  # noqa: E402
from lib.network.rtpose_vgg import get_model  # noqa: E402
from lib.utils.common import draw_humans  # noqa: E402
from lib.utils.paf_to_pose import paf_to_pose_cpp  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cfg', default='./experiments/vgg19_368x368_sgd.yaml', type=str)
    ap.add_argument('--weight', type=str, default='')
    ap.add_argument('--image', type=str, default='')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--out', default='result.png')
    ap.add_argument('opts', default=None, nargs=argparse.REMAINDER)
    args = ap.parse_args()
    update_config(cfg, args)

    model = get_model('vgg19')
    if args.weight:
        model.load_state_dict(torch.load(args.weight))
    else:
        import _b200_alias, importlib
        arrays = importlib.import_module(_b200_alias.PKG + ".synthetic").he_state_arrays(1234)
        model.load_state_dict({k: torch.from_numpy(a) for k, a in zip(model.state_dict(), arrays)})
    model = torch.nn.DataParallel(model).cuda()
    model.float()
    model.eval()
    model.module.precision = args.precision

    img = cv2.imread(args.image) if args.image else np.random.RandomState(0).randint(0, 256, (368, 368, 3)).astype(np.uint8)
    with torch.no_grad():
        paf, heatmap, im_scale = get_outputs(img, model, 'rtpose')
    print('im_scale', im_scale, 'maps', heatmap.shape, paf.shape)
    humans = paf_to_pose_cpp(heatmap, paf, cfg)
    print('%d humans' % len(humans))
    cv2.imwrite(args.out, draw_humans(img, humans))


if __name__ == '__main__':
    main()
