cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q 2>&1 | tail -5
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench5.json 2> gpurun_out/r02_bench5.err; tail -2 gpurun_out/r02_bench5.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench5.json'))
print({k:d[k] for k in ['value','ms_per_step','gpu_launches','scaling']}, d['e2e']['value'], 'conv', d['roofline']['frac'], 'hist', d['roofline_hist']['frac'], 'blend', d['roofline_blend']['frac'], d['roofline_blend']['launch_ms'], d['roofline_blend']['moved_gbs'], d['roofline_blend']['pure_write_gbs'])
print(d['clocks'], d['config']['images_per_group'])
P
python scripts/dev_blend_probe.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"tta_blend_x2|paste_tiles" -s 4 -c 3 -o gpurun_out/r02_blend_fused python scripts/dev_blend_probe.py > gpurun_out/r02_blend_fused_ncu.log 2>&1; ls -la gpurun_out/r02_blend_fused.ncu-rep
