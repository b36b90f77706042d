set -x
cd $GRAFT_REPO_ROOT
python scripts/dev_hist_probe.py uniform network saturated trained flat > gpurun_out/r02_hist_probe4.log 2>&1; tail -6 gpurun_out/r02_hist_probe4.log
EDS_UPSAMPLE_TILED=0 python scripts/dev_concat_probe.py > gpurun_out/r02_concat_probe.log 2>&1
EDS_UPSAMPLE_TILED=1 python scripts/dev_concat_probe.py >> gpurun_out/r02_concat_probe.log 2>&1
cat gpurun_out/r02_concat_probe.log
python -m pytest tests/test_kernels_gpu.py -q -x -k "pr_hist or pr_scan" 2>&1 | tail -3
