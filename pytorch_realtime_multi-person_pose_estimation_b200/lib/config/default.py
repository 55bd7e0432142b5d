"""Configuration for the inference path.  Mirrors the keys of /root/reference/lib/config/default.py that the hot
path reads (DATASET.IMAGE_SIZE :69, MODEL.DOWNSAMPLE :41, MODEL.NUM_KEYPOINTS :40, TEST.THRESH_HEATMAP :126,
TEST.THRESH_PAF :127, TEST.NUM_INTERMED_PTS_BETWEEN_KEYPOINTS :128) and `update_config(cfg, args)` (:139-168).
yacs is not available offline, so a small attribute-dict node with the same merge calls is used."""
import os

import yaml


class CfgNode(dict):
    """Minimal stand-in for yacs.config.CfgNode: attribute access, merge_from_file/list, freeze/defrost."""

    def __init__(self, init=None, new_allowed=True):
        super().__init__()
        self.__dict__["_frozen"] = False
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        if self.__dict__.get("_frozen"):
            raise AttributeError("config is frozen")
        self[k] = v

    def defrost(self):
        self.__dict__["_frozen"] = False
        for v in self.values():
            if isinstance(v, CfgNode):
                v.defrost()

    def freeze(self):
        self.__dict__["_frozen"] = True
        for v in self.values():
            if isinstance(v, CfgNode):
                v.freeze()

    def _merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict):
                if not isinstance(self.get(k), CfgNode):
                    self[k] = CfgNode()
                self[k]._merge(v)
            else:
                self[k] = v

    def merge_from_file(self, path):
        with open(path) as f:
            self._merge(yaml.safe_load(f) or {})

    def merge_from_list(self, opts):
        if not opts:
            return
        if len(opts) % 2:
            raise ValueError("opts must be KEY VALUE pairs")
        for key, val in zip(opts[0::2], opts[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = yaml.safe_load(val) if isinstance(val, str) else val

    def clone(self):
        return CfgNode(self)


_C = CfgNode({
    "OUTPUT_DIR": "", "LOG_DIR": "", "EXPERIMENT_NAME": "", "DATA_DIR": "", "GPUS": [0], "WORKERS": 4,
    "MODEL": {"NAME": "rtpose_vgg19", "NUM_KEYPOINTS": 18, "DOWNSAMPLE": 8, "HEATMAP_SIZE": [46, 46], "SIGMA": 7},
    "DATASET": {"ROOT": "", "DATASET": "coco", "IMAGE_SIZE": 368, "VAL_IMAGE_DIR": "", "VAL_ANNOTATIONS": ""},
    "TEST": {"BATCH_SIZE_PER_GPU": 32, "FLIP_TEST": False, "THRESH_HEATMAP": 0.1, "THRESH_PAF": 0.05,
             "NUM_INTERMED_PTS_BETWEEN_KEYPOINTS": 10, "MODEL_FILE": ""},
})


def update_config(cfg, args):
    cfg.defrost()
    if getattr(args, "cfg", None):
        cfg.merge_from_file(args.cfg)
    cfg.merge_from_list(getattr(args, "opts", None))
    for name in ("modelDir", "logDir", "dataDir"):
        val = getattr(args, name, None)
        if val:
            {"modelDir": lambda v: cfg.__setitem__("OUTPUT_DIR", v), "logDir": lambda v: cfg.__setitem__("LOG_DIR", v),
             "dataDir": lambda v: cfg.__setitem__("DATA_DIR", v)}[name](val)
    if cfg.get("DATA_DIR"):
        cfg.DATASET["ROOT"] = os.path.join(cfg.DATA_DIR, cfg.DATASET.get("ROOT", ""))
    cfg.freeze()
