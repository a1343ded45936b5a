cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/r02ac_bench.json 2> gpurun_out/r02ac_bench.err; echo rc=$?
python -c "
import json
d=json.load(open('gpurun_out/r02ac_bench.json'))
print('value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'pipe ms', d['e2e_pipelined']['ms_per_step'], d['result_check'])
print(d['phases_ms']); print('g2', d['g2']['ms_per_step'], 'ntt', d['ntt']['ms_per_step'], d['ntt']['roofline']['fr_mul_floor_frac'])
print('roof', d['roofline']['frac'], d['roofline']['executed_frac'])
"
python tools/gpu_profile_target.py msm24 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_msm_invert -c 14 --csv --log-file gpurun_out/r02ac_inv.csv python tools/gpu_profile_target.py msm24 > /dev/null 2>&1
grep invert gpurun_out/r02ac_inv.csv | awk -F'","' '{print $NF}' | tr -d '"' | tail -7 | tr '\n' ' '
