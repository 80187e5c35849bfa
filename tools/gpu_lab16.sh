#!/bin/bash
mkdir -p gpurun_out
{
for b in 128 64 32; do
  KG_K4_BLOCK=$b timeout 300 python tools/k4_ab.py --agents 1000000 --variants 0 --flush
  KG_K4_BLOCK=$b timeout 300 python tools/k4_ab.py --agents 1000000 --variants 0
done
KG_K4_BLOCK=64 timeout 300 python tools/k4_ab.py --agents 8000000 --variants 0 --steps 20
} > gpurun_out/lab16_ab.jsonl 2> gpurun_out/lab16_ab.err
cat gpurun_out/lab16_ab.jsonl; tail -3 gpurun_out/lab16_ab.err
