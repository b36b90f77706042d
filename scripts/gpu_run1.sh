cd $GRAFT_REPO_ROOT
python -m pytest tests/test_kernels_gpu.py -q 2>&1 | tail -40
