#!/bin/bash
# GPU visit 3: parity suite with the folded / phased plans, then A/B step timings of the new knobs.
TAG=${1:-r01_v4}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --tb=short 2>&1 | tail -120 > gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
B="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline"
echo "default:        $(timeout 300 $B 2>&1 | tail -1)"
echo "fold off:       $(CN_FOLD=0,1 timeout 300 $B 2>&1 | tail -1)"
echo "s2all off:      $(CN_FOLD=1,0 timeout 300 $B 2>&1 | tail -1)"
echo "chunk 8:        $(CN_CHUNK_KB=8 timeout 300 $B 2>&1 | tail -1)"
echo "chunk 4:        $(CN_CHUNK_KB=4 timeout 300 $B 2>&1 | tail -1)"
echo "cluster 1:      $(CN_CLUSTER=1 timeout 300 $B 2>&1 | tail -1)"
echo "cluster 4:      $(CN_CLUSTER=4 timeout 300 $B 2>&1 | tail -1)"
CN_CHUNK_KB=8 timeout 600 python -m pytest tests/test_stage2_gpu.py -m gpu -q --timeout 600 --tb=line -k "stage2_generator_step" 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/${TAG}_conv_breakdown.txt > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
