#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (N=1), ncu launch list and one full capture of the fused kernels.
# Usage (from the repo root):  gpurun --timeout 1500 -- 'bash profiles/gpu_round.sh <tag>'
tag=${1:-rXX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --batch 128 --e2e-batch 8 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(fwd|bwd)_fused' -s 4 -c 2 -f -o gpurun_out/${tag}_prof \
    python bench.py --steps 1 --warmup 3 --batch 128 --e2e-batch 8 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
timeout 300 python profiles/r50_probe.py 64 > gpurun_out/${tag}_r50.json 2> gpurun_out/${tag}_r50.err
timeout 300 python profiles/family_probe.py 256 > gpurun_out/${tag}_family.json 2> gpurun_out/${tag}_family.err
tail -3 gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_smoke.log | tail -2; cat gpurun_out/${tag}_bench_n1.json
