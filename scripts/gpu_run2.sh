cd $GRAFT_REPO_ROOT
EDS_TAG=plain python scripts/dev_concat_probe.py 2>&1 | tail -8
cp eyediseasesegmentation_b200/libeds_b200.so /tmp/plain.so; cp eyediseasesegmentation_b200/libeds_stream.so eyediseasesegmentation_b200/libeds_b200.so
EDS_TAG=stream python scripts/dev_concat_probe.py 2>&1 | tail -8
python scripts/dev_pass_probe.py 2>&1 | tail -1
cp /tmp/plain.so eyediseasesegmentation_b200/libeds_b200.so
python scripts/dev_pass_probe.py 2>&1 | tail -1
