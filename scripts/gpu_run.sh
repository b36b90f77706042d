cd $GRAFT_REPO_ROOT
python -m pytest tests/test_multigpu_gpu.py -q 2>&1 | tail -2
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_scale_${n}gpu.json 2> gpurun_out/r02_scale_${n}gpu.err
tail -1 gpurun_out/r02_scale_${n}gpu.err
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r02_scale_1gpu.json 2> gpurun_out/r02_scale_1gpu.err
python - <<'P'
import json
base=None
for n in (1,4,8):
    try:
        d=json.load(open(f'gpurun_out/r02_scale_{n}gpu.json'))
    except Exception as e:
        print(n, 'failed', e); continue
    if n==1: base=d['value']
    print(n, round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'ms/step', round(d['ms_per_step'],1), 'allreduce_ms', d['config']['allreduce_ms'], 'eff', round(d['value']/(n*base),3) if base else None, d['clocks'], d['gpu_launches'])
P
