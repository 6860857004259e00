#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (N=1), ncu launch list and full captures of the fused kernels.
# Usage (from the repo root):  gpurun --timeout 1800 -- 'bash profiles/gpu_round.sh <tag>'
tag=${1:-rXX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --batch 128 --level-graphs 4 --e2e-steps 1 --r50-batch 8 --e2e-op-batch 8 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(fwd|bwd)_fused' -s 8 -c 8 -f -o gpurun_out/${tag}_prof \
    python bench.py --steps 1 --warmup 3 --batch 128 --level-graphs 4 --e2e-steps 1 --r50-batch 8 --e2e-op-batch 8 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_smoke.log | tail -2; cat gpurun_out/${tag}_bench_n1.json
