#!/usr/bin/env python
"""Per-kernel SASS fingerprints of a built library: `python tools/sass_hash.py [lib.so] > a.txt`, diff two outputs to
prove that a change (a refactor, a new entry point, code behind a compile-time switch) left the machine code of the
kernels that were validated on the GPU untouched."""
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "pytorch_realtime_multi-person_pose_estimation_b200",
                                                         "libb200pose.so")
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
cur, table = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        table[cur] = hashlib.md5()
    elif cur and line.strip().startswith("/*"):      # instruction lines only
        table[cur].update(" ".join(line.split()).encode())      # column alignment varies with the widest line
for k in sorted(table):
    print(table[k].hexdigest()[:12], k)
