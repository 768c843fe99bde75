#!/bin/bash
TAG=${1:-r01_v6}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --tb=short 2>&1 | tail -60 > gpurun_out/${TAG}_pytest_gpu.log
tail -6 gpurun_out/${TAG}_pytest_gpu.log
B="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline"
echo "default:        $(timeout 300 $B 2>&1 | tail -1)"
timeout 300 python scripts/gpu_role_prof.py > gpurun_out/${TAG}_role_prof.txt 2>&1
grep -E "cluster=1|mma per|gather" gpurun_out/${TAG}_role_prof.txt | grep -A2 "cluster=1" | cut -c1-170
timeout 600 python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/${TAG}_conv_breakdown.txt > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json | cut -c1-400; tail -3 gpurun_out/${TAG}_bench.err
