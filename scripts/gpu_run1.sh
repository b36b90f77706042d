cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench10.json 2> gpurun_out/r02_bench10.err; tail -2 gpurun_out/r02_bench10.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench10.json'))
print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, d['e2e']['value'], 'conv', d['roofline']['frac'], 'hist', d['roofline_hist']['frac'], 'blend', d['roofline_blend']['frac'], d['clocks'])
print(d['time_shares'])
print(d['other_configs'])
P
