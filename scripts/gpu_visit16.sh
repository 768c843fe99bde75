cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_conv_bench_shapes_gpu.py -m gpu -q --tb=short 2>&1 | tail -6
echo "== TMA store epilogue (default)"; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --debug-steps --breakdown gpurun_out/r02_breakdown_tma.txt 2>gpurun_out/e.err | cut -c1-120; grep "value loop" gpurun_out/e.err | cut -c1-200
echo "== rule (no TMA store)"; CN_COAL=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --debug-steps --breakdown gpurun_out/r02_breakdown_notma.txt 2>gpurun_out/f.err | cut -c1-120; grep "value loop" gpurun_out/f.err | cut -c1-200
