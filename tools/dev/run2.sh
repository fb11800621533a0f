set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 180 --timeout-method=thread -k "golden_cases or fuzz_cuda_equals_oracle or packed or fused or solo or small_configs" 2>&1 | tail -5
timeout 300 python tools/dev/phase_prof.py cfg2-mod cfg2-random cfg-self 2>&1 | grep -v "^+" | tail -40
timeout 600 python bench.py --steps 30 --warmup 5 --sub cfg2-random,cfg-self --sub-steps 3 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -3 gpurun_out/bench_b.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_b.json'))
print('headline', d['value']/1e6, 'M/s', d['ms_per_step'], 'ms; e2e', d['e2e']['ms_per_step'], 'unpacked', (d.get('e2e_unpacked') or {}).get('ms_per_step'), 'parity', d['parity_in_run'])
for k,v in d['workloads'].items(): print(k, v['value']/1e6, v['ms_per_step'], 'e2e', v['e2e']['ms_per_step'], (v.get('e2e_unpacked') or {}).get('ms_per_step'), v['parity_in_run'], v['gpu_launches_per_step'])
print(d['dispatch_latency'])
P
