cd $GRAFT_REPO_ROOT
python tools/gpu_pipeline_probe.py 2
python tools/gpu_pipeline_probe.py 1 | tail -4
