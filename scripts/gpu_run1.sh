cd $GRAFT_REPO_ROOT
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r02_bench11.json 2> gpurun_out/r02_bench11.err; tail -3 gpurun_out/r02_bench11.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench11.json'))
print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, 'e2e', d['e2e']['value'], d['clocks'])
P
