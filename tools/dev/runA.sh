set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method=thread 2>&1 | tail -6
timeout 900 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; tail -3 gpurun_out/r2f_bench_n1.err
YDSCHED_FUSED_MAX_N=2097152 timeout 300 python bench.py --workload cfg3 --steps 5 --sub none --no-cpu-baseline --no-latency > gpurun_out/bench_cfg3_fused.json 2> gpurun_out/bench_cfg3_fused.err
python - <<'P'
import json
for f in ['gpurun_out/r2f_bench_n1.json','gpurun_out/bench_cfg3_fused.json']:
    try: d=json.load(open(f))
    except Exception as e: print(f, 'unreadable', e); continue
    print(f, 'headline', d['value']/1e6, 'M/s', d['ms_per_step'], 'ms; e2e', d['e2e']['ms_per_step'], 'unpacked', (d.get('e2e_unpacked') or {}).get('ms_per_step'), 'parity', d['parity_in_run'], d['roofline']['frac'], d['clocks'])
    for k,v in d['workloads'].items(): print(k, v['value']/1e6, v['ms_per_step'], 'e2e', v['e2e']['ms_per_step'], (v.get('e2e_unpacked') or {}).get('ms_per_step'), v['parity_in_run'], v['gpu_launches_per_step'])
    print(d['dispatch_latency'])
P
