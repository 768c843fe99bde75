#!/bin/bash
# GPU visit 2: full parity suite, skinny-layer timing, bench with per-layer table, step timeline.
TAG=${1:-r01_v3}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --tb=short 2>&1 | tail -300 > gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python -m pytest tests/test_stage2_gpu.py -m gpu -q --timeout 600 -k "stage2_generator_step or class_surface" 2>&1 | tail -250 > gpurun_out/${TAG}_pytest_stage2.log
timeout 300 python scripts/gpu_skinny_time.py > gpurun_out/${TAG}_skinny_time.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/${TAG}_conv_breakdown.txt > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 300 python scripts/gpu_step_timeline.py --steps 2 --out gpurun_out/${TAG}_timeline.txt > /dev/null 2>&1
tail -5 gpurun_out/${TAG}_pytest_gpu.log; tail -3 gpurun_out/${TAG}_pytest_stage2.log; cat gpurun_out/${TAG}_skinny_time.txt; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err; head -3 gpurun_out/${TAG}_timeline.txt
