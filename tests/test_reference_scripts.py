"""BASELINE.json north_star: "... so evaluate/evaluation.py and demo/picture_demo.py run unmodified".

CPU part (needs /root/reference, which exists in the build container only): the reference's OWN two scripts are read
from where they lie and executed, unmodified, against this repository's `lib.*` / `evaluate.*` modules, from the repo
root as the reference expects.  Nothing is copied.  Third-party modules the scripts import but never use on this path
(`pylab`, `matplotlib`, absent from the image) are stubbed, `torch.load` hands out a synthetic checkpoint (no
checkpoint exists offline), and execution is stopped at the first `.cuda()` - the point where a GPU becomes necessary.
Everything before it (every import, argparse + update_config on the shipped yaml, get_model, load_state_dict with the
reference's key names, DataParallel) has then run for real, and the calls after it are checked to bind.

GPU part: this repo's demo with the reference's command line (no extra flags), on the shipped 674x712 stand-in for
./readme/ski.jpg and a pose_model.pth written by the test."""
import inspect
import os
import subprocess
import sys
import types

import pytest
import torch

from conftest import ROOT

REF = "/root/reference"


class _ReachedCuda(Exception):
    pass


def _exec_reference_script(rel, monkeypatch, fake_load):
    path = os.path.join(REF, rel)
    src = open(path).read()
    monkeypatch.chdir(ROOT)
    monkeypatch.setattr(sys, "argv", [os.path.basename(rel)])
    for name in ("pylab", "matplotlib"):
        try:
            __import__(name)
        except ImportError:
            monkeypatch.setitem(sys.modules, name, types.ModuleType(name))
    monkeypatch.setattr(torch, "load", fake_load)

    def cuda(self, *a, **k):
        raise _ReachedCuda(type(self).__name__)
    monkeypatch.setattr(torch.nn.Module, "cuda", cuda)
    glob = {"__name__": "__main__", "__file__": path}
    with pytest.raises(_ReachedCuda) as ei:
        exec(compile(src, path, "exec"), glob)
    return glob, str(ei.value)


needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree exists in the build container only")


@needs_ref
def test_reference_picture_demo_runs_unmodified_up_to_cuda(monkeypatch, built, he_sd):
    glob, who = _exec_reference_script("demo/picture_demo.py", monkeypatch, lambda *a, **k: he_sd)
    assert who == "DataParallel"                       # model = torch.nn.DataParallel(model).cuda()   (:47)
    model = glob["model"]                              # still the bare module: the assignment never completed
    assert len(model.state_dict()) == 184 and torch.equal(model.state_dict()["model0.0.weight"], he_sd["model0.0.weight"])
    assert glob["cfg"].DATASET.IMAGE_SIZE == 368 and glob["args"].weight == "pose_model.pth"
    assert os.path.exists(os.path.join(ROOT, "readme", "ski.jpg"))        # test_image = './readme/ski.jpg'   (:51)
    # the calls the script makes after .cuda() bind to this repo's functions (:58-64)
    inspect.signature(glob["get_outputs"]).bind("oriImg", "model", "rtpose")
    inspect.signature(glob["paf_to_pose_cpp"]).bind("heatmap", "paf", glob["cfg"])
    inspect.signature(glob["draw_humans"]).bind("oriImg", [])
    inspect.signature(glob["handle_paf_and_heat"]).bind(1, 2, 3, 4)
    for name in ("Human", "BodyPart", "CocoPart", "CocoColors", "CocoPairsRender", "im_transform", "update_config"):
        assert name in glob


@needs_ref
def test_reference_evaluation_script_runs_unmodified_up_to_cuda(monkeypatch, built, he_sd):
    ckpt = {"state_dict": {"model." + k: v for k, v in he_sd.items()}}     # evaluation.py:15-20 strips 6 characters
    glob, who = _exec_reference_script("evaluate/evaluation.py", monkeypatch, lambda *a, **k: ckpt)
    assert who == "rtpose_model"                       # model = model.cuda()   (:27)
    assert list(glob["new_state_dict"]) == list(he_sd)
    assert torch.equal(glob["model"].state_dict()["model6_2.12.bias"], he_sd["model6_2.12.bias"])
    inspect.signature(glob["run_eval"]).bind(image_dir="a", anno_file="b", vis_dir="c", model=glob["model"], preprocess="vgg")
    assert "OpenPose_Model" in glob and "use_vgg" in glob


@pytest.mark.gpu
def test_demo_with_the_reference_command_line(built, he_sd):
    """`python demo/picture_demo.py` from the repo root, no flags: reads ./experiments/vgg19_368x368_sgd.yaml,
    pose_model.pth and ./readme/ski.jpg, writes result.png - the reference demo's contract (:30-65)."""
    weight, result = os.path.join(ROOT, "pose_model.pth"), os.path.join(ROOT, "result.png")
    torch.save(he_sd, weight)
    try:
        if os.path.exists(result):
            os.remove(result)
        r = subprocess.run([sys.executable, os.path.join("demo", "picture_demo.py")], cwd=ROOT, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True, timeout=600)
        print(r.stdout[-2000:])
        assert r.returncode == 0 and os.path.exists(result)
        assert abs(float(r.stdout.split()[0]) - 368.0 / 674.0) < 1e-12        # print(im_scale)
        assert "humans; maps (46, 49, 19) (46, 49, 38)" in r.stdout           # 674x712 -> 368x389 -> padded 368x392
    finally:
        for f in (weight, result):
            if os.path.exists(f):
                os.remove(f)
