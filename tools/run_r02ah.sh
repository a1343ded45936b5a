cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_msm.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r02ah_pytest.log
cat gpurun_out/r02ah_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02ah_bench.json 2> gpurun_out/r02ah_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02ah_bench.json').read().strip().splitlines()[-1])
print('value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'pipelined', d['e2e_pipelined']['ms_per_step'], d['result_check'])
PY
python tools/gpu_sweep_factor.py 21 1,2,4,8,16 2>&1 | tee gpurun_out/r02ah_factor.txt
python tools/gpu_sweep_factor.py 24 1,2,4 2>&1 | tee -a gpurun_out/r02ah_factor.txt
