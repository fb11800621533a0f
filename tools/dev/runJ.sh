cd $GRAFT_REPO_ROOT
exec > gpurun_out/runJ.log 2>&1
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 60 --timeout-method=thread -k "test_c_example_against_cuda_library" 2>&1 | tail -4
