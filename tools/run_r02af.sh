cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r02af_pytest.log 2>&1; tail -4 gpurun_out/r02af_pytest.log | head -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 > gpurun_out/r02af_bench.json 2> gpurun_out/r02af_bench.err; echo rc=$? lines=$(wc -l < gpurun_out/r02af_bench.json)
python -c "
import json
d=json.load(open('gpurun_out/r02af_bench.json'))
print('value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'pipe ms', d['e2e_pipelined']['ms_per_step'], d['result_check'])
print('g2', d['g2']['ms_per_step'], d['g2']['roofline'])
print('ntt', d['ntt']['ms_per_step'], d['ntt']['e2e']['ms_per_step'])
print([ (x['log_n'], round(x['ours_ms'],2)) for x in d['reference_gpu']['g1_msm']['sizes']])
"
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | head -c 300
