set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
python scripts/dev_hist_probe.py > gpurun_out/r02_hist_probe.log 2>&1; tail -8 gpurun_out/r02_hist_probe.log
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu1.log 2>&1; tail -15 gpurun_out/r02_gpu1.log
python __graft_entry__.py smoke > gpurun_out/r02_smoke1.log 2>&1; tail -3 gpurun_out/r02_smoke1.log
