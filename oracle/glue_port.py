"""ORACLE (test infrastructure only).  CPU restatement of the inference glue of the reference:

  * crop_with_factor / _factor_closest   /root/reference/lib/network/im_transform.py:113-134
  * rtpose_preprocess / vgg_preprocess   /root/reference/lib/datasets/preprocessing.py:16-21, 32-43
  * get_outputs                          /root/reference/evaluate/coco_eval.py:80-114
  * handle_paf_and_heat                  /root/reference/evaluate/coco_eval.py:197-242
  * resize_cubic / multi_scale_maps      composition for BASELINE.json configs[4] (cv2.resize INTER_CUBIC restated)
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


def resize_cubic(src, dh, dw):
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_CUBIC) on float32 [h,w] / [h,w,C] maps, restated from OpenCV's
    published algorithm exactly like nms_port.upsample8_cubic but for any size ratio: f = (float)((d+0.5)*(src/dst)-0.5),
    taps floor(f)-1..floor(f)+2 with replicated borders, interpolateCubic(A=-0.75) in float32, horizontal pass
    ((s0*a0+s1*a1)+s2*a2)+s3*a3, vertical pass s0*b0+(s1*b1+(s2*b2+s3*b3)).  cv2 itself evaluates the last
    (row_length mod 4) elements of a row in another order (SIMD tail), so this matches cv2 (IPP off) only to 2.4e-7,
    and the IPP-enabled default of this container to 2e-5 (tests/test_oracle.py)."""
    F32 = np.float32
    src = np.ascontiguousarray(src, dtype=F32)
    squeeze = src.ndim == 2
    if squeeze:
        src = src[:, :, None]
    h, w, _ = src.shape

    def axis(dn, sn):
        step = 1.0 / (float(dn) / float(sn))
        f = ((np.arange(dn, dtype=np.float64) + 0.5) * step - 0.5).astype(F32)
        s = np.floor(f).astype(np.int64)
        co = np.stack([nms_port._cubic_coeffs(x) for x in (f - s.astype(F32)).astype(F32)]).astype(F32)
        return np.clip(s[:, None] - 1 + np.arange(4)[None, :], 0, sn - 1), co
    ix, ax = axis(dw, w)
    iy, by = axis(dh, h)
    hor = src[:, ix[:, 0], :] * ax[None, :, 0, None]
    for j in (1, 2, 3):
        hor = hor + src[:, ix[:, j], :] * ax[None, :, j, None]
    out = hor[iy[:, 3]] * by[:, 3, None, None]
    for j in (2, 1, 0):
        out = hor[iy[:, j]] * by[:, j, None, None] + out
    out = out.astype(F32)
    return out[:, :, 0] if squeeze else out


def multi_scale_maps(per_scale, base_hw):
    """Multi-scale test-time averaging (BASELINE.json configs[4]; the reference at this commit has no multi-scale loop,
    so this is a composition of its functions): per_scale = [(heat, paf)] or [(heat, paf, heat_flipped, paf_flipped)]
    per scale, HWC float32 at that scale's grid.  Each scale is flip-merged with handle_paf_and_heat when the mirrored
    maps are given, resized to base_hw with resize_cubic, summed in float32 in the given order and divided by the
    number of scales.  Returns (avg_heat, avg_paf)."""
    F32 = np.float32
    acc_h = acc_p = None
    for item in per_scale:
        heat, paf = item[0], item[1]
        if len(item) == 4:
            paf, heat = handle_paf_and_heat(heat, item[2], paf, item[3])
            heat, paf = heat.astype(F32), paf.astype(F32)
        rh, rp = resize_cubic(heat, *base_hw), resize_cubic(paf, *base_hw)
        acc_h = rh if acc_h is None else (acc_h + rh).astype(F32)
        acc_p = rp if acc_p is None else (acc_p + rp).astype(F32)
    n = F32(len(per_scale))
    return (acc_h / n).astype(F32), (acc_p / n).astype(F32)


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
