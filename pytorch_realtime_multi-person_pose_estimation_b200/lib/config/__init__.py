"""`lib.config` of the reference exposes two names (demo/picture_demo.py:26, evaluate/coco_eval.py:17):
the shared configuration node `cfg` and `update_config(cfg, args)`.  Both live in `default.py`."""
from . import default as _default

cfg = _default._C
update_config = _default.update_config

__all__ = ["cfg", "update_config"]
