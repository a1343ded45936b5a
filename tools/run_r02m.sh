cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
set -x
python tools/gpu_profile_target.py ntt24 > gpurun_out/r02m_plain_ntt.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_ntt -s 3 -c 3 -f -o gpurun_out/r02m_ntt python tools/gpu_profile_target.py ntt24 > gpurun_out/r02m_ncu_ntt.log 2>&1
python tools/gpu_profile_target.py msm24 > gpurun_out/r02m_plain_msm.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_msm_hist|k_msm_scatter|k_msm_pair_fwd|k_msm_pair_bwd|k_msm_invert' -s 23 -c 11 -f -o gpurun_out/r02m_msm python tools/gpu_profile_target.py msm24 > gpurun_out/r02m_ncu_msm.log 2>&1
python tools/gpu_profile_target.py g2_20 > gpurun_out/r02m_plain_g2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_msm_pair_fwd|k_msm_pair_bwd|k_msm_accumulate|k_msm_segment|k_msm_combine' -s 10 -c 10 -f -o gpurun_out/r02m_g2 python tools/gpu_profile_target.py g2_20 > gpurun_out/r02m_ncu_g2.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02m_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02m_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02m_ncu_bench.log 2>&1
ls -la gpurun_out/r02m_*
tail -3 gpurun_out/r02m_ncu_ntt.log gpurun_out/r02m_ncu_msm.log gpurun_out/r02m_ncu_g2.log gpurun_out/r02m_ncu_bench.log
