set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu4.log 2>&1; tail -4 gpurun_out/r02_gpu4.log
python bench.py --steps 4 --warmup 3 > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err; tail -3 gpurun_out/r02_bench3.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench3.json'))
print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, d['e2e']['value'])
print('conv',d['roofline']['frac'],'hist', d['roofline_hist']['frac'],'blend', d['roofline_blend']['frac'], d['roofline_blend']['launch_ms'], d['roofline_blend']['merge_only'])
print(d['other_configs'])
print(d['cpu_baseline'])
P
python scripts/dev_hist_probe.py uniform > gpurun_out/hp.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pr_hist_kernel -s 1 -c 2 -o gpurun_out/r02_hist python scripts/dev_hist_probe.py uniform > gpurun_out/r02_hist_ncu.log 2>&1
python scripts/dev_blend_probe.py > gpurun_out/bp.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"tta_merge64|paste_tiles" -s 8 -c 4 -o gpurun_out/r02_blend python scripts/dev_blend_probe.py > gpurun_out/r02_blend_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r02_bench_under_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_bench_launches.csv | tail -5
