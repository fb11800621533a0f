set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 180 --timeout-method=thread -k "golden_cases or fuzz_cuda_equals_oracle or packed or fused or solo or small_configs or many_classes or prefiltered or full_size" 2>&1 | tail -4
timeout 200 python tools/dev/phase_prof.py cfg2-mod 2>&1 | grep -v "^+" | grep -A1 "staged cold" | tail -6
# merge chunk sweep
for c in 128 256 512 1024; do for w in cfg2-random cfg-self; do echo "chunk $c $w: $(YDSCHED_MERGE_CHUNK=$c YDSCHED_MERGE_ROUNDS=64 timeout 120 python bench.py --workload $w --steps 10 --sub none --no-cpu-baseline --no-latency 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"]*1e3,1), "us value; e2e", round(d["e2e"]["ms_per_step"]*1e3,1))')"; done; done
for c in 128 256; do echo "chunk $c cfg3: $(YDSCHED_MERGE_CHUNK=$c YDSCHED_MERGE_ROUNDS=64 timeout 120 python bench.py --workload cfg3 --steps 5 --sub none --no-cpu-baseline --no-latency 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"]*1e3,1), "us value")')"; done
# memcheck over the new kernels (small cases)
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_fuzz_packed_interface or test_solo_kernel_stands_down or (test_fused_front_and_packed and small) or (prefiltered and both)" > gpurun_out/r2f_sanitizer_memcheck.log 2>&1; echo memcheck rc=$?; tail -4 gpurun_out/r2f_sanitizer_memcheck.log
timeout 600 python bench.py --steps 100 --warmup 5 --sub none > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_f.json'))
print('headline', d['value']/1e6, 'M/s', d['ms_per_step'], 'ms; e2e', d['e2e']['ms_per_step'], d['e2e_ms_steps'], 'unpacked', (d.get('e2e_unpacked') or {}).get('ms_per_step'), 'parity', d['parity_in_run'])
print(d['dispatch_latency'])
P
