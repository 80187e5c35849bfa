#!/usr/bin/env python
"""Which byte values make the fused Forest-Fire passes and K5 disagree?  Both against a numpy model that
only looks at bits 0 and 1 of a cell (diagnostic for tests/test_gpu_grid.py's arbitrary-bytes case)."""
import numpy as np
import krabmaga_b200 as kb


def model_step(v):
    low = v & 3
    burning = low == 2
    green = low == 1
    p = np.pad(burning, 1)
    fire = np.zeros_like(burning)
    for dx in (0, 1, 2):
        for dy in (0, 1, 2):
            fire |= p[dx:dx + v.shape[0], dy:dy + v.shape[1]]
    return (v + ((green & fire) | burning).astype(np.uint8)).astype(np.uint8)


def main():
    w, h = 150, 1984
    rng = np.random.default_rng(11)
    cells = rng.integers(0, 256, size=(w, h), dtype=np.uint8)
    for name, steps in (("single", 1), ("single", 14), ("fused2", 2), ("fused8", 8), ("fused14", 14)):
        g = kb.DenseNumberGrid2D(w, h)
        g.upload(cells, unbuffered=True)
        g.lazy_update()
        if name == "single":
            for _ in range(steps):
                g.step_stencil()
                g.lazy_update()
        else:
            g.run_stencil(steps)
        got = g.download()
        want = cells.copy()
        for _ in range(steps):
            want = model_step(want)
        bad = np.argwhere(got != want)
        print(name, "mismatches", len(bad))
        if len(bad):
            vals = sorted(set(int(cells[x, y]) for x, y in bad[:2000]))
            print("  original bytes at mismatching cells:", [hex(v) for v in vals][:40])
            for x, y in bad[:8]:
                print("  cell", x, y, "orig", hex(int(cells[x, y])), "got", hex(int(got[x, y])), "want", hex(int(want[x, y])),
                      "y%512", y % 512, "y%960", y % 960)


if __name__ == "__main__":
    main()
