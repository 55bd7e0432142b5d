"""ORACLE (test infrastructure only).  CPU restatement of the inference glue of the reference:

  * crop_with_factor / _factor_closest   /root/reference/lib/network/im_transform.py:113-134
  * rtpose_preprocess / vgg_preprocess   /root/reference/lib/datasets/preprocessing.py:16-21, 32-43
  * get_outputs                          /root/reference/evaluate/coco_eval.py:80-114
  * handle_paf_and_heat                  /root/reference/evaluate/coco_eval.py:197-242
  * paf_to_pose_cpp                      /root/reference/lib/utils/paf_to_pose.py:372-406

cv2.resize (bilinear, uint8) is the reference's own third-party call and is used the same way here.
"""
import cv2
import numpy as np
import torch

from . import net_port, nms_port

SWAP_HEAT = np.array((0, 1, 5, 6, 7, 2, 3, 4, 11, 12, 13, 8, 9, 10, 15, 14, 17, 16, 18))       # coco_eval.py:207
SWAP_PAF = np.array((6, 7, 8, 9, 10, 11, 0, 1, 2, 3, 4, 5, 20, 21, 22, 23, 24, 25, 26, 27, 12, 13, 14, 15, 16, 17,
                     18, 19, 28, 29, 32, 33, 30, 31, 36, 37, 34, 35))                           # coco_eval.py:228


def crop_with_factor(im, dest_size, factor=8):
    scale = float(dest_size) / min(im.shape[0:2])
    im = cv2.resize(im, None, fx=scale, fy=scale)
    h, w, c = im.shape
    nh = int(np.ceil(float(h) / factor)) * factor
    nw = int(np.ceil(float(w) / factor)) * factor
    out = np.zeros([nh, nw, c], dtype=im.dtype)
    out[0:h, 0:w, :] = im
    return out, scale, im.shape


def rtpose_preprocess(image):
    image = image.astype(np.float32) / 256. - 0.5
    return image.transpose((2, 0, 1)).astype(np.float32)


def vgg_preprocess(image):
    image = image.astype(np.float32) / 255.
    out = image.copy()[:, :, ::-1]
    for i, (m, s) in enumerate(zip([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])):
        out[:, :, i] = out[:, :, i] - m
        out[:, :, i] = out[:, :, i] / s
    return out.transpose((2, 0, 1)).astype(np.float32)


def get_outputs(img, sd, preprocess="rtpose", inp_size=368, factor=8):
    im, scale, _ = crop_with_factor(img, inp_size, factor)
    data = rtpose_preprocess(im) if preprocess == "rtpose" else vgg_preprocess(im)
    with torch.no_grad():
        (paf, heat), _ = net_port.forward(sd, torch.from_numpy(data[None]))
    return paf.numpy().transpose(0, 2, 3, 1)[0], heat.numpy().transpose(0, 2, 3, 1)[0], scale


def handle_paf_and_heat(normal_heat, flipped_heat, normal_paf, flipped_paf):
    fp = flipped_paf[:, ::-1, :].copy()
    fp[:, :, SWAP_PAF[::2]] = -fp[:, :, SWAP_PAF[::2]]     # negate x components (all even channels)
    avg_paf = (normal_paf + fp[:, :, SWAP_PAF]) / 2.
    avg_heat = (normal_heat + flipped_heat[:, ::-1, :][:, :, SWAP_HEAT]) / 2.
    return avg_paf, avg_heat


def paf_to_pose(heat, paf, pafprocess, thresh=0.1, upsample=8, num_keypoints=18):
    """paf_to_pose_cpp with `pafprocess` = any object exposing the SWIG module's functions.
    Returns (joint_list [P,5] float32, humans [(score, {part: (x/W, y/H, peak_score)})])."""
    per_joint = nms_port.nms(heat, thresh, num_keypoints)
    jl = nms_port.joint_list_from_nms(per_joint)
    humans = []
    if jl.shape[0] > 0:
        paf_up = np.repeat(np.repeat(paf, upsample, axis=0), upsample, axis=1)      # == cv2 INTER_NEAREST x8
        heat_up = np.repeat(np.repeat(heat, upsample, axis=0), upsample, axis=1)
        pafprocess.process_paf(jl[None], np.ascontiguousarray(heat_up), np.ascontiguousarray(paf_up))
        H, W = heat_up.shape[:2]
        for hid in range(pafprocess.get_num_humans()):
            parts = {}
            for p in range(num_keypoints):
                c = int(pafprocess.get_part_cid(hid, p))
                if c < 0:
                    continue
                parts[p] = (float(pafprocess.get_part_x(c)) / W, float(pafprocess.get_part_y(c)) / H,
                            float(pafprocess.get_part_score(c)))
            if parts:
                humans.append((float(pafprocess.get_score(hid)), parts))
    return jl, humans
