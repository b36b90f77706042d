set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_multi.log 2>&1; tail -6 gpurun_out/r02_gpu_multi.log
python scripts/dev_hist_probe.py uniform network saturated trained flat > gpurun_out/r02_hist_probe5.log 2>&1; tail -6 gpurun_out/r02_hist_probe5.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 4 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench_2gpu.json'))
print({k:d[k] for k in ['value','ms_per_step','n_gpus','steps','scaling','gpu_launches']}, d['e2e']['value'], d['config']['allreduce_ms'])
P
tail -5 gpurun_out/r02_bench_2gpu.err
python bench.py --steps 6 --warmup 4 --no-cpu-baseline --no-extra > gpurun_out/r02_bench_1gpu_s6.json 2> gpurun_out/r02_bench_1gpu_s6.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench_1gpu_s6.json'))
print({k:d[k] for k in ['value','ms_per_step','n_gpus','steps','scaling','gpu_launches']}, d['e2e']['value'], d['roofline_hist']['frac'], d['roofline_blend']['frac'])
P
