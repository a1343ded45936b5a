cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
set -x
(time python bench.py --steps 5 --warmup 3) > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err
echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r02q_bench.json'))
print('value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'pipe ms', d['e2e_pipelined']['ms_per_step'], d['result_check'])
"
cap() {  # name, regex, skip, count, target
  python tools/gpu_profile_target.py $5 > gpurun_out/r02q_plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o /tmp/r02q_$1 python tools/gpu_profile_target.py $5 > gpurun_out/r02q_ncu_$1.log 2>&1
  ncu -i /tmp/r02q_$1.ncu-rep --page raw --csv > gpurun_out/r02q_$1_raw.csv 2>/dev/null
  ls -la /tmp/r02q_$1.ncu-rep
}
cap ntt 'k_ntt' 3 3 ntt24
cap msm 'k_msm_hist|k_msm_scatter|k_msm_pair_fwd|k_msm_pair_bwd|k_msm_invert' 23 11 msm24
cap g2 'k_msm_pair_fwd|k_msm_pair_bwd|k_msm_accumulate|k_msm_segment|k_msm_combine' 10 10 g2_20
cap tail21 'k_msm_accumulate|k_msm_finalize|k_msm_segment|k_msm_tree|k_msm_combine|k_msm_task|k_scan' 0 40 msm21
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02q_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02q_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02q_ncu_bench.log 2>&1
du -sh gpurun_out
