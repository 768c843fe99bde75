#!/bin/bash
TAG=${1:-r01_v11}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_networks_gpu.py -m gpu -q --timeout 300 --tb=short -k "graph_replayed" 2>&1 | tail -40 > gpurun_out/${TAG}_pytest_graph.log
tail -12 gpurun_out/${TAG}_pytest_graph.log
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --tb=short 2>&1 | tail -40 > gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/${TAG}_conv_breakdown.txt > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-1000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
echo "graphs off: $(CN_GRAPHS=0 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-700)"
