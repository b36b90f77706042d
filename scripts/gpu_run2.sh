cd $GRAFT_REPO_ROOT
for i in 1 2; do
EDS_SE_EPILOGUE=0 python scripts/dev_pass_probe.py 2>&1 | tail -1
EDS_SE_EPILOGUE=1 EDS_SE_EPILOGUE_MIN_PIXELS=999999999 python scripts/dev_pass_probe.py 2>&1 | tail -1
EDS_SE_EPILOGUE=1 python scripts/dev_pass_probe.py 2>&1 | tail -1
EDS_SE_EPILOGUE=1 EDS_SE_EPILOGUE_MIN_PIXELS=1 python scripts/dev_pass_probe.py 2>&1 | tail -1
done
