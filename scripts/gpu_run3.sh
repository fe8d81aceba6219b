python -m pytest tests -m gpu -x -q 2>&1 | tail -8
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
$B > gpurun_out/r2_b3.json 2> gpurun_out/r2_b3.err; tail -c 800 gpurun_out/r2_b3.err
$B --ms-flags 4 > gpurun_out/r2_b3_nopair.json 2>> gpurun_out/r2_b3.err
$B --ms-flags 8 > gpurun_out/r2_b3_unfused.json 2>> gpurun_out/r2_b3.err
$B --chunk-len 48 > gpurun_out/r2_b3_c48.json 2>> gpurun_out/r2_b3.err
$B --chunk-len 96 > gpurun_out/r2_b3_c96.json 2>> gpurun_out/r2_b3.err
$B --chunk-len 128 > gpurun_out/r2_b3_c128.json 2>> gpurun_out/r2_b3.err
$B --e2e-depth 1 > gpurun_out/r2_b3_d1.json 2>> gpurun_out/r2_b3.err
$B --e2e-depth 2 > gpurun_out/r2_b3_d2.json 2>> gpurun_out/r2_b3.err
$B --e2e-threads 6 > gpurun_out/r2_b3_t6.json 2>> gpurun_out/r2_b3.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b3*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']; print(f, 'value', round(d['value']/1e9,1), 'single', round(d['impl_detail']['single_stream']['value']/1e9,1), 'e2e', round(d['e2e']['value']/1e9,1), r['kernel_ms']['pack'], r['kernel_ms']['ms_fused'], r['events_per_base'], round(r['algorithmic_bytes_per_base'],1), 'frac', round(r['frac'],3), r['bound'])
    except Exception as ex: print(f, 'ERR', ex)
PY
