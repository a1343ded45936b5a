cd $GRAFT_REPO_ROOT
python tools/gpu_profile_target.py msm24 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_msm|k_scan|k_copy' --csv --log-file gpurun_out/r02ay_msm24_launches.csv python tools/gpu_profile_target.py msm24 > gpurun_out/r02ay_ncu.log 2>&1
wc -l gpurun_out/r02ay_msm24_launches.csv
