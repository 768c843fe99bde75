#!/bin/bash
# Round-end style visit: parity suite, smoke, both bench arms, ncu launch list (+DRAM bytes) of one bench
# invocation, ncu --set full of the tensor-core conv kernel.
TAG=${1:-r01_final}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --tb=short 2>&1 | tail -30 > gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/${TAG}_conv_breakdown.txt > gpurun_out/${TAG}_bench.json 2>> gpurun_out/${TAG}_bench.err
cut -c1-600 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
CN_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --roofline-pass inline > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm_tc -c 8 -f -o gpurun_out/${TAG}_tc_full \
    python scripts/ncu_conv_one.py > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
timeout 300 python scripts/gpu_step_timeline.py --steps 2 --out gpurun_out/${TAG}_timeline.txt > /dev/null 2>&1
head -3 gpurun_out/${TAG}_timeline.txt | cut -c1-200
