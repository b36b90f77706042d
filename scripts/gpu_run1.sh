cd $GRAFT_REPO_ROOT
python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_golden_gpu.py tests/test_configs_gpu.py -q -x 2>&1 | tail -4
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench9.json 2> gpurun_out/r02_bench9.err; tail -2 gpurun_out/r02_bench9.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench9.json'))
print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, d['e2e']['value'], 'conv', d['roofline']['frac'], d['clocks'])
print(d['time_shares'])
print(d['other_configs']['cfg2_uppse50_bf16_16x1024_flip'])
P
