cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 5 > gpurun_out/r02_bench_2gpu_b.json 2> gpurun_out/r02_bench_2gpu_b.err; tail -2 gpurun_out/r02_bench_2gpu_b.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench_2gpu_b.json'))
print({k:d[k] for k in ['value','ms_per_step','n_gpus','gpu_launches']}, 'e2e', d['e2e']['value'], d['config']['allreduce_ms'])
P
python -m pytest tests/test_multigpu_gpu.py -q 2>&1 | tail -1
