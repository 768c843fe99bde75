#!/bin/bash
# quick A/B visit: conv operator parity, step time, per-layer dbg timings
TAG=${1:-quick}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 120 --tb=short -x 2>&1 | tail -30 > gpurun_out/${TAG}_pytest_ops.log
tail -3 gpurun_out/${TAG}_pytest_ops.log
if ! grep -q " passed" gpurun_out/${TAG}_pytest_ops.log || grep -q "failed\|Timeout\|error" gpurun_out/${TAG}_pytest_ops.log; then echo "OPS TESTS NOT CLEAN - stopping"; exit 0; fi
B="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline"
echo "default:     $(timeout 300 $B 2>&1 | tail -1)"
timeout 300 python scripts/gpu_role_prof.py > gpurun_out/${TAG}_role_prof.txt 2>&1
grep "dbg=" gpurun_out/${TAG}_role_prof.txt | sed -E 's/cluster=1 dbg=([0-9]+) \(16 = no epilogue stores, 4 = no MMA issue\): ([0-9.]+) us\/call.*/dbg=\1 \2us/'
