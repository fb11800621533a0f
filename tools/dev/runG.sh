cd $GRAFT_REPO_ROOT
exec > gpurun_out/runG.log 2>&1
echo "== A: host prof, packed + plain + staged"
YDSCHED_HOST_PROF=1 timeout 300 python bench.py --workload cfg5 --steps 3 --warmup 3 --sub none --no-cpu-baseline --no-latency 2>&1 | grep -v "^{" | tail -30
echo "== B: plain call alone"
BENCH_NO_PACKED=1 timeout 300 python bench.py --workload cfg5 --steps 3 --warmup 3 --sub none --no-cpu-baseline --no-latency | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("plain alone: value", d["ms_per_step"], "e2e", d["e2e"])'
nvidia-smi topo -m | head -12
numactl --hardware 2>/dev/null | head -8
