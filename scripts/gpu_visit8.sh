cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "== run A (first process on the box)"; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --debug-steps --breakdown gpurun_out/r02_breakdown_lpr2_rule.txt 2>gpurun_out/a.err | cut -c1-200; grep loop gpurun_out/a.err | cut -c1-400
echo "== run B (same)"; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --debug-steps 2>gpurun_out/b.err | cut -c1-200; grep loop gpurun_out/b.err | cut -c1-400
echo "== coal never"; CN_COAL=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --breakdown gpurun_out/r02_breakdown_lpr2_coal0.txt 2>/dev/null | cut -c1-200
echo "== coal always"; CN_COAL=2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --breakdown gpurun_out/r02_breakdown_lpr2_coal2.txt 2>/dev/null | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q --timeout 900 --tb=short 2>&1 | tail -5
