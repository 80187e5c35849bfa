#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/blocks_scale_probe.py 32000000 2>&1 | tail -3
timeout 300 python tools/k4_ab.py --agents 32000000 --variants 0 --steps 10 --settle 5 2>&1 | tail -2
