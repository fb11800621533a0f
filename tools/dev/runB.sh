set -x
cd $GRAFT_REPO_ROOT
timeout 900 python tools/ncu_traffic.py cfg2-mod cfg2-random cfg-self > gpurun_out/r2f_dram_traffic.json 2> gpurun_out/ncu_traffic.err; tail -3 gpurun_out/ncu_traffic.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_front -s 2 -c 1 -o gpurun_out/r2f_fused_front_cfg2mod -f python tools/one_solve.py cfg2-mod 4 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
timeout 300 python bench.py --workload cfg4 --steps 20 --sub none --no-latency > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_cfg4.json'))
print('cfg4', d['value']/1e6, d['ms_per_step'], 'e2e', d['e2e'], d['parity_in_run'])
t=json.load(open('gpurun_out/r2f_dram_traffic.json'))
for k,v in t['workloads'].items(): print(k, {a:b for a,b in v.items() if not a.startswith('launches')})
P
ls -la gpurun_out | tail -20
