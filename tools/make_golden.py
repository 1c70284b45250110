"""Regenerate tests/golden/oracle_checksums.json and tests/golden/readme_60x30.npz from the CPU
oracle.  The reference (Julia) cannot run here, so the golden vectors are (a) the README's
printed numbers and the zipper test's exact expectations, encoded directly in
tests/test_oracle_golden.py, and (b) these bit-level pins of the oracle that reproduced them."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as orc          # noqa: E402
from test_oracle_golden import _checksums  # noqa: E402

o = orc.Oracle()
os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
json.dump(_checksums(o), open(os.path.join(ROOT, "tests", "golden", "oracle_checksums.json"), "w"), indent=1, sort_keys=True)
p = orc.make_params((60, 30, 1))
g = o.build(p)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "readme_60x30.npz"), **g)
print("golden files written")
