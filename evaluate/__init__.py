"""`evaluate.*` (the reference's import path for coco_eval) -> the package's evaluate/* modules."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _b200_alias  # noqa: E402

_b200_alias.install("evaluate", "evaluate")
