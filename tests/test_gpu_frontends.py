"""GPU (-m gpu): the rows of SURVEY.md 8(f) ranks 3 and 4 - the batched evaluation loop and the streaming front-ends -
against the reference-shaped per-image sequence (get_outputs + paf_to_pose_cpp) they replace."""
import json
import os
import subprocess
import sys
import types

import numpy as np
import pytest
import torch

from conftest import ROOT, pkg_module

pytestmark = pytest.mark.gpu


def _model(he_sd, precision="bf16"):
    from lib.network.rtpose_vgg import get_model
    model = get_model("vgg19")
    model.load_state_dict(he_sd)
    model = model.cuda().float().eval()
    model.precision = precision
    return model


class Opaque(torch.nn.Module):
    """Hides the device fast paths: coco_eval then runs the reference's own per-image sequence through the same net."""

    def __init__(self, inner):
        super().__init__()
        self.inner = inner

    def forward(self, x):
        return self.inner(x)


def test_batched_run_eval_equals_the_per_image_loop(built, he_sd, monkeypatch, tmp_path):
    """run_eval over a directory of images of mixed sizes: the batched fused path (shape buckets, EVAL_BATCH frames per
    pass) must produce the very same COCO records and visualisation files as get_outputs + paf_to_pose_cpp per image."""
    import cv2
    ev = pkg_module("evaluate.coco_eval")
    img_dir = tmp_path / "img"
    img_dir.mkdir()
    rs = np.random.RandomState(5)
    shapes = [(120, 160), (200, 230), (120, 160), (368, 368), (200, 230), (97, 64), (120, 160)]
    names = ["im%d.png" % i for i in range(len(shapes))]          # png: lossless, both runs read identical pixels
    for name, (h, w) in zip(names, shapes):
        cv2.imwrite(str(img_dir / name), rs.randint(0, 256, (h, w, 3)).astype(np.uint8))
    seen = {}

    class COCO:
        def __init__(self, anno): pass
        def getCatIds(self, catNms): return [1]
        def getImgIds(self, catIds): return list(range(len(names)))
        def loadImgs(self, i): return [{"file_name": names[i]}]
        def loadRes(self, path): seen["results"] = json.load(open(path)); return "dt"

    class COCOeval:
        def __init__(self, gt, dt, kind): self.params = types.SimpleNamespace(imgIds=None); self.stats = [0.25]
        def evaluate(self): pass
        def accumulate(self): pass
        def summarize(self): pass
    for name, attrs in (("pycocotools", {}), ("pycocotools.coco", {"COCO": COCO}), ("pycocotools.cocoeval", {"COCOeval": COCOeval})):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        monkeypatch.setitem(sys.modules, name, m)
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(ev, "EVAL_BATCH", 4)          # several passes, several buckets per pass
    model = _model(he_sd)
    out = {}
    for tag, mdl in (("batched", model), ("serial", Opaque(model))):
        vis = tmp_path / ("vis_" + tag)
        vis.mkdir()
        with torch.no_grad():
            assert ev.run_eval(str(img_dir), "anno.json", str(vis), model=mdl, preprocess="vgg") == 0.25
        out[tag] = (seen["results"], {n: cv2.imread(str(vis / n)) for n in names})
    assert len(out["batched"][0]) > 50
    assert out["batched"][0] == out["serial"][0]
    for n in names:
        np.testing.assert_array_equal(out["batched"][1][n], out["serial"][1][n])
    print("run_eval: %d COCO records over %d images, batched == per-image" % (len(out["batched"][0]), len(names)))


def test_pose_stream_equals_the_per_frame_loop(built, he_sd):
    """streaming.PoseStream (batches of 3, two in flight) vs the reference front-end's loop body
    (get_outputs + paf_to_pose_cpp + draw_humans per frame, web_demo.py:55-66)."""
    from evaluate.coco_eval import get_outputs
    from lib.config import cfg
    from lib.utils.common import draw_humans
    from lib.utils.paf_to_pose import paf_to_pose_cpp
    streaming = pkg_module("streaming")
    model = _model(he_sd)
    rs = np.random.RandomState(11)
    frames = [rs.randint(0, 256, (240, 320, 3)).astype(np.uint8) for _ in range(7)]
    got = list(streaming.PoseStream(model, frames, batch=3))
    assert len(got) == 7
    for frame, (f2, humans, drawn) in zip(frames, got):
        assert f2 is frame
        with torch.no_grad():
            paf, heat, _ = get_outputs(frame, model, 'rtpose')
        want = paf_to_pose_cpp(heat, paf, cfg)
        assert len(humans) == len(want) > 0
        for a, b in zip(humans, want):
            assert a.score == b.score and a.body_parts.keys() == b.body_parts.keys()
            for k in a.body_parts:
                pa, pb = a.body_parts[k], b.body_parts[k]
                assert (pa.x, pa.y, pa.score, pa.uidx) == (pb.x, pb.y, pb.score, pb.uidx)
        np.testing.assert_array_equal(drawn, draw_humans(frame, want, imgcopy=True))


def test_video_and_camera_front_ends_run(built, tmp_path):
    """video_demo.py on a synthetic clip (the reference's flow: every frame drawn into an output video) and
    demo/web_demo.py headless on the same clip as its "camera"."""
    import cv2
    clip = str(tmp_path / "clip.avi")
    w = cv2.VideoWriter(clip, cv2.VideoWriter_fourcc(*'MJPG'), 20.0, (320, 240))
    rs = np.random.RandomState(3)
    for _ in range(11):
        w.write(rs.randint(0, 256, (240, 320, 3)).astype(np.uint8))
    w.release()
    out = str(tmp_path / "output.avi")
    r = subprocess.run([sys.executable, "video_demo.py", "--synthetic-weights", "--video", clip, "--out", out, "--batch", "4"],
                       cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    print(r.stdout[-1500:])
    assert r.returncode == 0 and "11 frames" in r.stdout and "Shape of image is  (320, 240)" in r.stdout
    cap = cv2.VideoCapture(out)
    n = 0
    while cap.read()[0]:
        n += 1
    assert n == 11
    r = subprocess.run([sys.executable, os.path.join("demo", "web_demo.py"), "--synthetic-weights", "--source", clip,
                        "--no-window", "--max-frames", "6"], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=600)
    print(r.stdout[-800:])
    assert r.returncode == 0 and "6 frames" in r.stdout
