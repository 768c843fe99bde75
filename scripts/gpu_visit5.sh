cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_conv_bench_shapes_gpu.py -m gpu -q --tb=short 2>&1 | tail -8
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/r02_breakdown_cg_coal.txt > gpurun_out/r02_bench_cg_coal.json 2> gpurun_out/r02_bench_cg_coal.err; cut -c1-260 gpurun_out/r02_bench_cg_coal.json; tail -3 gpurun_out/r02_bench_cg_coal.err
CN_DBG=32 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/r02_breakdown_cg_nocoal.txt > gpurun_out/r02_bench_cg_nocoal.json 2> gpurun_out/r02_bench_cg_nocoal.err; cut -c1-260 gpurun_out/r02_bench_cg_nocoal.json; tail -3 gpurun_out/r02_bench_cg_nocoal.err
