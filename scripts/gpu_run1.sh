cd $GRAFT_REPO_ROOT
EDS_CONCAT_SKIP_LEAN=0 python scripts/dev_concat_probe.py 2>&1 | tail -8
EDS_CONCAT_SKIP_LEAN=1 python scripts/dev_concat_probe.py 2>&1 | tail -8
python -m pytest tests/test_kernels_gpu.py -q -x -k "deferred_gate or concat_gated or gaussian" 2>&1 | tail -2
