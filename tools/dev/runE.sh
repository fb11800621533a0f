set -x
cd $GRAFT_REPO_ROOT
exec > gpurun_out/runE.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method=thread 2>&1 | tail -6
timeout 900 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; tail -3 gpurun_out/r2f_bench_n1.err
timeout 600 python tools/ncu_traffic.py cfg2-mod cfg2-random cfg-self > gpurun_out/r2f_dram_traffic.json 2> gpurun_out/ncu_traffic.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2f_bench_n1.json'))
print('headline', d['value']/1e6, 'M/s', d['ms_per_step'], 'ms; e2e', d['e2e']['ms_per_step'], d['e2e_ms_steps'], 'unpacked', (d.get('e2e_unpacked') or {}).get('ms_per_step'), 'parity', d['parity_in_run'])
for k,v in d['workloads'].items(): print(k, v['value']/1e6, v['ms_per_step'], 'e2e', v['e2e']['ms_per_step'], (v.get('e2e_unpacked') or {}).get('ms_per_step'), v['parity_in_run'], v['gpu_launches_per_step'])
print(d['dispatch_latency'])
P
