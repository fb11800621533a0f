cd $GRAFT_REPO_ROOT
exec > gpurun_out/runI.log 2>&1
timeout 300 python -m pytest tests/test_service.py -m gpu -x -q --timeout 120 --timeout-method=thread 2>&1 | tail -5
