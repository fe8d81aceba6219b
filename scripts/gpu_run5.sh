python -m pytest tests -m gpu -x -q 2>&1 | tail -6
B="python bench.py --steps 20 --warmup 5"
$B > gpurun_out/r2_b5.json 2> gpurun_out/r2_b5.err; tail -c 800 gpurun_out/r2_b5.err
$B --no-cpu-baseline --ms-flags 16 > gpurun_out/r2_b5_fused.json 2>> gpurun_out/r2_b5.err
$B --no-cpu-baseline --e2e-depth 2 > gpurun_out/r2_b5_d2.json 2>> gpurun_out/r2_b5.err
$B --no-cpu-baseline --e2e-depth 4 > gpurun_out/r2_b5_d4.json 2>> gpurun_out/r2_b5.err
$B --no-cpu-baseline --e2e-threads 6 > gpurun_out/r2_b5_t6.json 2>> gpurun_out/r2_b5.err
$B --no-cpu-baseline --steps 200 > gpurun_out/r2_b5_s200.json 2>> gpurun_out/r2_b5.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b5*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']; print(f, 'value', round(d['value']/1e9,1), 'single', round(d['impl_detail']['single_stream']['value']/1e9,1), 'e2e', round(d['e2e']['value']/1e9,1), r['kernel_ms']['pack'], r['kernel_ms']['ms_fused'], 'frac', round(r['frac'],3), r['bound'], d['gpu_launches'])
    except Exception as ex: print(f, 'ERR', ex)
PY
