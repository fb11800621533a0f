cd $GRAFT_REPO_ROOT
exec > gpurun_out/runF.log 2>&1
timeout 900 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; tail -3 gpurun_out/r2f_bench_n1.err
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 180 --timeout-method=thread -k "packed or fused or solo or small_configs" 2>&1 | tail -3
python - <<'P'
import json
d=json.load(open('gpurun_out/r2f_bench_n1.json'))
print('headline', d['value']/1e6, 'M/s', d['ms_per_step'], 'ms; e2e', d['e2e']['ms_per_step'], d['e2e_ms_steps'], 'unpacked', (d.get('e2e_unpacked') or {}).get('ms_per_step'), 'parity', d['parity_in_run'])
for k,v in d['workloads'].items(): print(k, v['value']/1e6, v['ms_per_step'], 'e2e', v['e2e']['ms_per_step'], (v.get('e2e_unpacked') or {}).get('ms_per_step'), v['parity_in_run'], v['gpu_launches_per_step'])
print(d['dispatch_latency'])
P
