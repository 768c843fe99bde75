#!/bin/bash
TAG=${1:-r01_v8}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 120 --tb=short -x 2>&1 | tail -30 > gpurun_out/${TAG}_pytest_ops.log
tail -5 gpurun_out/${TAG}_pytest_ops.log
if ! grep -q " passed" gpurun_out/${TAG}_pytest_ops.log || grep -q "failed\|Timeout\|error" gpurun_out/${TAG}_pytest_ops.log; then echo "OPS TESTS NOT CLEAN - stopping"; exit 0; fi
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --tb=short 2>&1 | tail -60 > gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
B="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline"
echo "persistent:     $(timeout 300 $B 2>&1 | tail -1)"
echo "non-persistent: $(CN_PERSISTENT=0 timeout 300 $B 2>&1 | tail -1)"
timeout 300 python scripts/gpu_role_prof.py > gpurun_out/${TAG}_role_prof.txt 2>&1
grep "dbg=0" gpurun_out/${TAG}_role_prof.txt | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/${TAG}_conv_breakdown.txt > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-300 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
