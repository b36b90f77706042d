cd $GRAFT_REPO_ROOT
python scripts/dev_se_probe.py 2>&1 | tail -4
