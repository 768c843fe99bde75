#!/bin/bash
TAG=${1:-r01_v5}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --tb=short 2>&1 | tail -60 > gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
B="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline"
echo "default:        $(timeout 300 $B 2>&1 | tail -1)"
echo "wcache off:     $(CN_WCACHE=0 timeout 300 $B 2>&1 | tail -1)"
echo "cluster 1:      $(CN_CLUSTER=1 timeout 300 $B 2>&1 | tail -1)"
timeout 600 python scripts/gpu_cluster_ab.py profiles/r01_conv_breakdown_v4.txt > gpurun_out/${TAG}_cluster_ab.txt 2>&1
tail -3 gpurun_out/${TAG}_cluster_ab.txt
timeout 300 python scripts/gpu_step_timeline.py --steps 2 --out gpurun_out/${TAG}_timeline.txt > /dev/null 2>&1
head -30 gpurun_out/${TAG}_timeline.txt | cut -c1-140
