"""Drop-in for the inference glue of /root/reference/evaluate/coco_eval.py: `get_outputs` (:80-114) and
`handle_paf_and_heat` (:197-242).  COCO mAP scoring (eval_coco / run_eval with pycocotools) is outside the hot
path and not provided.  Unlike the reference this module does not parse sys.argv at import;
`cfg` is the shared lib.config node."""
import numpy as np
import torch

from ..lib.config import cfg
from ..lib.datasets.preprocessing import (inception_preprocess, rtpose_preprocess, ssd_preprocess, vgg_preprocess)
from ..lib.network import im_transform

ORDER_COCO = [0, 15, 14, 17, 16, 5, 2, 6, 3, 7, 4, 11, 8, 12, 9, 13, 10]   # coco_eval.py:52


def get_outputs(img, model, preprocess):
    """img: HWC uint8 BGR -> (paf [h,w,38] f32, heatmap [h,w,19] f32, im_scale)."""
    im_croped, im_scale, _ = im_transform.crop_with_factor(img, cfg.DATASET.IMAGE_SIZE, factor=cfg.MODEL.DOWNSAMPLE,
                                                           is_ceil=True)
    if preprocess == 'rtpose':
        im_data = rtpose_preprocess(im_croped)
    elif preprocess == 'vgg':
        im_data = vgg_preprocess(im_croped)
    elif preprocess == 'inception':
        im_data = inception_preprocess(im_croped)
    elif preprocess == 'ssd':
        im_data = ssd_preprocess(im_croped)
    batch_var = torch.from_numpy(np.expand_dims(im_data, 0)).cuda().float()   # unknown mode -> UnboundLocalError, as upstream
    predicted_outputs, _ = model(batch_var)
    output1, output2 = predicted_outputs[-2], predicted_outputs[-1]
    heatmap = output2.cpu().data.numpy().transpose(0, 2, 3, 1)[0]
    paf = output1.cpu().data.numpy().transpose(0, 2, 3, 1)[0]
    return paf, heatmap, im_scale


_SWAP_HEAT = np.array((0, 1, 5, 6, 7, 2, 3, 4, 11, 12, 13, 8, 9, 10, 15, 14, 17, 16, 18))
_SWAP_PAF = np.array((6, 7, 8, 9, 10, 11, 0, 1, 2, 3, 4, 5, 20, 21, 22, 23, 24, 25, 26, 27, 12, 13, 14, 15, 16, 17, 18,
                      19, 28, 29, 32, 33, 30, 31, 36, 37, 34, 35))


def handle_paf_and_heat(normal_heat, flipped_heat, normal_paf, flipped_paf):
    """Left/right-flip test-time averaging: mirror W, negate the PAF x components, swap left/right channels."""
    mirrored = flipped_paf[:, ::-1, :]
    mirrored[:, :, _SWAP_PAF[::2]] = -mirrored[:, :, _SWAP_PAF[::2]]   # in place on the caller's array, as upstream
    averaged_paf = (normal_paf + mirrored[:, :, _SWAP_PAF]) / 2.
    averaged_heatmap = (normal_heat + flipped_heat[:, ::-1, :][:, :, _SWAP_HEAT]) / 2.
    return averaged_paf, averaged_heatmap
