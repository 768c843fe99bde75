#!/bin/bash
# Round-2 closing check on the final tree: full parity suite, smoke, the headline bench line and config 1 (complete lines).
TAG=${1:-r02_m9}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 400 --tb=short > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log | cut -c1-200
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --breakdown gpurun_out/${TAG}_conv_breakdown.txt > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cut -c1-250 gpurun_out/${TAG}_bench.json
timeout 200 python bench.py --config 1 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_cfg1.json 2> gpurun_out/${TAG}_bench_cfg1.err; cut -c1-250 gpurun_out/${TAG}_bench_cfg1.json
