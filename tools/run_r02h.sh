cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_points.py -x -q) > gpurun_out/r02h_pytest.log 2>&1
tail -n 15 gpurun_out/r02h_pytest.log
which compute-sanitizer; ls /usr/local/cuda/bin | grep -i sanit; compute-sanitizer --version 2>&1 | head -3
