#!/bin/bash
# One GPU visit: parity tests, smoke, bench (+ per-layer table), reference arm, ncu launch list, ncu --set full
# of the tensor-core conv kernels.  Everything lands in gpurun_out/<tag>_*.
TAG=${1:-r01_v2}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
lscpu | head -20 >> gpurun_out/${TAG}_smi.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -40 > gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/${TAG}_conv_breakdown.txt > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm_tc -c 8 -f -o gpurun_out/${TAG}_tc_full \
    python scripts/ncu_conv_one.py > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_smoke.log | tail -3; cat gpurun_out/${TAG}_bench.json; cat gpurun_out/${TAG}_bench_reference.json
