cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r02an_pytest.log 2>&1; tail -3 gpurun_out/r02an_pytest.log
(time python bench.py --steps 5 --warmup 3) > gpurun_out/r02an_bench.json 2> gpurun_out/r02an_bench.err
echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02an_bench_reference.json 2> gpurun_out/r02an_bench_reference.err
echo "ref rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02an_bench.json').read().strip().splitlines()[-1])
print('value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'pipe ms', d['e2e_pipelined']['ms_per_step'], d['result_check'], 'ntt', d['ntt']['ms_per_step'], d['ntt'].get('result_check'), 'g2', d['g2']['ms_per_step'])
"
cap() {  # name, regex, skip, count, target
  python tools/gpu_profile_target.py $5 > gpurun_out/r02an_plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o /tmp/r02an_$1 python tools/gpu_profile_target.py $5 > gpurun_out/r02an_ncu_$1.log 2>&1
  ncu -i /tmp/r02an_$1.ncu-rep --page raw --csv > gpurun_out/r02an_$1_raw.csv 2>/dev/null
  ls -la /tmp/r02an_$1.ncu-rep
}
cap msm 'k_msm_hist|k_msm_scatter|k_msm_pair_fwd|k_msm_pair_bwd|k_msm_invert' 23 11 msm24
cap ntt 'k_ntt' 3 3 ntt24
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02an_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02an_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02an_ncu_bench.log 2>&1
du -sh gpurun_out
