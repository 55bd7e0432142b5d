"""ORACLE (test infrastructure only - never imported by the product path).

CPU fp32 restatement of the rtpose VGG19 network of the reference:
  * layer table     /root/reference/lib/network/rtpose_vgg.py:69-83 (trunk), :95-105 (stage 1), :108-127 (stages 2-6)
  * forward / concat /root/reference/lib/network/rtpose_vgg.py:158-198
  * "last conv of a branch has no ReLU"  rtpose_vgg.py:30-35

The convolution arithmetic itself is the reference's third-party dependency (torch, requirements.txt:2
`torch>=1.2`; container 2.11.0+cu128, CPU = oneDNN) and is called the same way (F.conv2d / F.max_pool2d /
torch.cat).  Pinned against the imported reference module in tests/golden (tests/golden/make_golden.py).
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# (name, cin, cout, k) ; 'P' = MaxPool2d(2, 2, 0).  Sequential indices follow conv,relu,(pool) numbering.
TRUNK = [("conv1_1", 3, 64, 3), ("conv1_2", 64, 64, 3), "P",
         ("conv2_1", 64, 128, 3), ("conv2_2", 128, 128, 3), "P",
         ("conv3_1", 128, 256, 3), ("conv3_2", 256, 256, 3), ("conv3_3", 256, 256, 3), ("conv3_4", 256, 256, 3), "P",
         ("conv4_1", 256, 512, 3), ("conv4_2", 512, 512, 3), ("conv4_3_CPM", 512, 256, 3), ("conv4_4_CPM", 256, 128, 3)]
NUM_STAGES = 6
PAF_CH, HEAT_CH, FEAT_CH = 38, 19, 128


def stage_layers(stage, branch):
    """[(cin, cout, k)] for model{stage}_{branch}; the last one has no ReLU."""
    out_ch = PAF_CH if branch == 1 else HEAT_CH
    if stage == 1:
        return [(128, 128, 3)] * 3 + [(128, 512, 1), (512, out_ch, 1)]
    cin = PAF_CH + HEAT_CH + FEAT_CH
    return [(cin, 128, 7)] + [(128, 128, 7)] * 4 + [(128, 128, 1), (128, out_ch, 1)]


def state_dict_spec():
    """OrderedDict key -> shape, in the reference module's state_dict order (184 tensors): model0 first, then
    model1_1..model6_1, then model1_2..model6_2 (attribute order of rtpose_model.__init__, rtpose_vgg.py:140-156)."""
    spec = OrderedDict()
    idx = 0
    for item in TRUNK:
        if item == "P":
            idx += 1
            continue
        _, cin, cout, k = item
        spec["model0.%d.weight" % idx] = (cout, cin, k, k)
        spec["model0.%d.bias" % idx] = (cout,)
        idx += 2
    for branch in (1, 2):
        for stage in range(1, NUM_STAGES + 1):
            for li, (cin, cout, k) in enumerate(stage_layers(stage, branch)):
                spec["model%d_%d.%d.weight" % (stage, branch, 2 * li)] = (cout, cin, k, k)
                spec["model%d_%d.%d.bias" % (stage, branch, 2 * li)] = (cout,)
    return spec


def he_state_dict(seed=1234):
    """Seeded variance-preserving synthetic weights: W ~ N(0, 2/fan_in), b ~ U(-0.1, 0.1).
    (The reference's own init, std=0.01 at rtpose_vgg.py:200-206, gives outputs ~1e-10: useless for parity.)"""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for key, shape in state_dict_spec().items():
        if key.endswith("weight"):
            fan_in = shape[1] * shape[2] * shape[3]
            sd[key] = torch.randn(shape, generator=g) * float(np.sqrt(2.0 / fan_in))
        else:
            sd[key] = torch.rand(shape, generator=g) * 0.2 - 0.1
    return sd


def forward(sd, x):
    """x: float32 [N,3,H,W] -> ((paf, heat), saved_for_loss[12]) exactly as rtpose_vgg.py:158-198."""
    idx = 0
    for item in TRUNK:
        if item == "P":
            x = F.max_pool2d(x, 2, 2, 0)
            idx += 1
            continue
        x = F.relu(F.conv2d(x, sd["model0.%d.weight" % idx], sd["model0.%d.bias" % idx], padding=item[3] // 2))
        idx += 2
    feat = x
    saved = []
    inp = feat
    for stage in range(1, NUM_STAGES + 1):
        outs = []
        for branch in (1, 2):
            y = inp
            layers = stage_layers(stage, branch)
            for li, (_, _, k) in enumerate(layers):
                y = F.conv2d(y, sd["model%d_%d.%d.weight" % (stage, branch, 2 * li)],
                             sd["model%d_%d.%d.bias" % (stage, branch, 2 * li)], padding=k // 2)
                if li != len(layers) - 1:
                    y = F.relu(y)
            outs.append(y)
        saved += outs
        inp = torch.cat([outs[0], outs[1], feat], 1)
    return (saved[-2], saved[-1]), saved
