"""Generates tests/golden/refnumpy_config1_20steps.npz from tests/ref_numpy.py (the numpy.float32
restatement that shares no code with oracle/): BASELINE config 1, State::init positions and the
state after 20 steps in canonical bag order.  Run from the repo root:
    python tests/golden/make_refnumpy_golden.py
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ref_numpy as rn  # noqa: E402

DISC = float(np.float32(10.0) / np.float32(1.5))
t0 = time.time()
wd = rn.World(400.0, 400.0, DISC, True, seed=42)
wd.init(10000)
first = wd.by_id()
wd.step(20)
last = wd.by_id()
np.savez_compressed(os.path.join(HERE, "refnumpy_config1_20steps.npz"), x0=first["x"], y0=first["y"],
                    x=last["x"], y=last["y"], ldx=last["ldx"], ldy=last["ldy"])
print(f"wrote refnumpy_config1_20steps.npz in {time.time() - t0:.0f} s")
