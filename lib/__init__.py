"""`lib.*` (the reference's import paths used by demo/picture_demo.py and evaluate/evaluation.py) -> the package's
lib/* modules.  See _b200_alias.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _b200_alias  # noqa: E402

_b200_alias.install("lib", "lib")
