cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 200 python scripts/gpu_tf32_peak.py > gpurun_out/r02_tf32_peak.txt 2>&1; tail -20 gpurun_out/r02_tf32_peak.txt
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_conv_bench_shapes_gpu.py -m gpu -q --tb=short 2>&1 | tail -8
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/r02_breakdown_coal.txt > gpurun_out/r02_bench_coal.json 2> gpurun_out/r02_bench_coal.err; cut -c1-260 gpurun_out/r02_bench_coal.json
CN_DBG=32 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/r02_breakdown_nocoal.txt > gpurun_out/r02_bench_nocoal.json 2> gpurun_out/r02_bench_nocoal.err; cut -c1-260 gpurun_out/r02_bench_nocoal.json
timeout 400 python bench.py --config 4 --steps 6 --warmup 3 --no-cpu-baseline --debug-steps > gpurun_out/r02_bench_cfg4_dbg.json 2> gpurun_out/r02_bench_cfg4_dbg.err; cut -c1-260 gpurun_out/r02_bench_cfg4_dbg.json; cat gpurun_out/r02_bench_cfg4_dbg.err | cut -c1-400
