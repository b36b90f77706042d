cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q 2>&1 | tail -2
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/plain_s1.json 2> gpurun_out/plain_s1.err && ncu --metrics gpu__time_duration.sum --clock-control none -s 4300 -c 1050 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r02_bench_under_ncu.log 2>&1
wc -l gpurun_out/r02_bench_launches.csv
