#!/bin/bash
mkdir -p gpurun_out
timeout 800 python tools/blocks_scale_probe.py 64000000 > gpurun_out/lab36_blocks_n8.txt 2>&1; cat gpurun_out/lab36_blocks_n8.txt | tail -5
