cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_points.py tests/test_gpu_msm.py -x -q -m gpu 2>&1 | tail -2
python tools/gpu_sweep_c.py 24 16 d | head -1
python tools/gpu_sweep_c.py 21 16 d | head -1
python tools/gpu_profile_target.py msm24 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_msm_invert -c 14 --csv --log-file gpurun_out/r02ag_inv.csv python tools/gpu_profile_target.py msm24 > /dev/null 2>&1
grep invert gpurun_out/r02ag_inv.csv | awk -F'","' '{print $NF}' | tr -d '"' | tail -7 | tr '\n' ' '
