"""Drop-in for the image normalisations of /root/reference/lib/datasets/preprocessing.py:16-86 (HWC uint8 BGR ->
CHW float32)."""
import numpy as np


def rtpose_preprocess(image):
    return (image.astype(np.float32) / 256. - 0.5).transpose((2, 0, 1)).astype(np.float32)


def inverse_rtpose_preprocess(image):
    return ((image.astype(np.float32).transpose((1, 2, 0)) + 0.5) * 256.).astype(np.uint8)


def vgg_preprocess(image):
    rgb = (image.astype(np.float32) / 255.)[:, :, ::-1].copy()
    for i, (m, s) in enumerate(zip((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))):
        rgb[:, :, i] = (rgb[:, :, i] - m) / s
    return rgb.transpose((2, 0, 1)).astype(np.float32)


def inception_preprocess(image):
    rgb = image.copy()[:, :, ::-1].astype(np.float32)
    return (rgb / 128. - 1.).transpose((2, 0, 1)).astype(np.float32)


def ssd_preprocess(image):
    # preprocessing.py:75-86: subtract (104, 117, 123) from (R, G, B), result kept in B, G, R order
    img = image.astype(np.float32).copy()
    img -= np.array((123.0, 117.0, 104.0), dtype=np.float32)
    return img.transpose((2, 0, 1)).astype(np.float32)
