#!/usr/bin/env python
"""Rows-per-tile sweep of the T-step Forest-Fire kernel on a W x H grid of one GPU (the shape one strip
of a multi-GPU run owns): prints ms per step for KG_FFT_ROWS / KG_FF_FUSE combinations (each in its own
process: the library reads the hooks once).
usage: python tools/ff_tile_probe.py W H"""
import os
import subprocess
import sys

CHILD = r"""
import sys, krabmaga_b200 as kb
w, h = int(sys.argv[1]), int(sys.argv[2])
g = kb.DenseNumberGrid2D(w, h)
g.init_forest_fire(0.6, 42)
g.run_stencil(16)
ms = g.run_stencil_timed(400)
print(ms / 400)
"""

def main():
    w, h = sys.argv[1], sys.argv[2]
    for fuse in ("8", "4"):
        for rows in ("0", "32", "48", "64", "96", "128"):
            env = dict(os.environ, KG_FF_FUSE=fuse)
            if rows != "0":
                env["KG_FFT_ROWS"] = rows
            out = subprocess.run([sys.executable, "-c", CHILD, w, h], env=env, capture_output=True, text=True)
            print(f"{w}x{h} T<={fuse} rows={'auto' if rows == '0' else rows}: {out.stdout.strip() or out.stderr[-300:]} ms/step", flush=True)

if __name__ == "__main__":
    main()
