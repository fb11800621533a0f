set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method=thread 2>&1 | tail -6
timeout 300 python tools/dev/phase_prof.py cfg2-mod 2>&1 | grep -v "^+" | tail -8
timeout 600 python bench.py --steps 30 --warmup 5 --sub cfg2-random,cfg-self,cfg3 --sub-steps 3 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; tail -3 gpurun_out/bench_c.err
YDSCHED_NO_ZEROCOPY=1 timeout 600 python bench.py --steps 30 --warmup 5 --sub none --no-cpu-baseline --no-latency > gpurun_out/bench_c_nozc.json 2> gpurun_out/bench_c_nozc.err
python - <<'P'
import json
for f in ['gpurun_out/bench_c.json','gpurun_out/bench_c_nozc.json']:
    d=json.load(open(f))
    print(f, 'headline', d['value']/1e6, 'M/s', d['ms_per_step'], 'ms; e2e', d['e2e']['ms_per_step'], 'unpacked', (d.get('e2e_unpacked') or {}).get('ms_per_step'), 'parity', d['parity_in_run'])
    for k,v in d['workloads'].items(): print(k, v['value']/1e6, v['ms_per_step'], 'e2e', v['e2e']['ms_per_step'], (v.get('e2e_unpacked') or {}).get('ms_per_step'), v['parity_in_run'], v['gpu_launches_per_step'])
    print(d['dispatch_latency'])
P
