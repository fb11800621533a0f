cd $GRAFT_REPO_ROOT
exec > gpurun_out/runH.log 2>&1
timeout 600 python bench.py --steps 20 --sub cfg5,cfg4 --sub-steps 3 > gpurun_out/bench_final_check.json 2> gpurun_out/bench_final_check.err; echo rc=$?; tail -3 gpurun_out/bench_final_check.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_final_check.json'))
print('headline', d['value']/1e6, d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'parity', d['parity_in_run'])
for k,v in d['workloads'].items(): print(k, v['value']/1e6, v['ms_per_step'], 'e2e', v['e2e']['ms_per_step'], (v.get('e2e_unpacked') or {}).get('ms_per_step'), v['parity_in_run'])
P
python -c "import __graft_entry__ as g; g.smoke()"
