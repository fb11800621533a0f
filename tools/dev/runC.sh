set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py > gpurun_out/r2f_multi_gpu_parity_2gpu.log 2>&1; echo rc=$?; tail -12 gpurun_out/r2f_multi_gpu_parity_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err; echo rc=$?; tail -3 gpurun_out/r2f_bench_n2.err; cut -c1-900 gpurun_out/r2f_bench_n2.json
