#!/bin/bash
# ncu capture of the tensor-core feature-mix kernels (forward, grad-X, grad-W) and the R50 kernels.
# Usage: gpurun --timeout 1200 -- 'bash profiles/gpu_mix_ncu.sh <tag>'
tag=${1:-rXX}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_mix_(fwd|gx|gw)_tc' -s 6 -c 3 -f -o gpurun_out/${tag}_mix_prof \
    python profiles/mix_probe.py big > gpurun_out/${tag}_mix_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_mix_launches.csv \
    python profiles/mix_probe.py big > gpurun_out/${tag}_mix_ncu2.log 2>&1
tail -3 gpurun_out/${tag}_mix_ncu.log
