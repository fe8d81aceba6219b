B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
$B > gpurun_out/r2_b11.json 2> gpurun_out/r2_b11.err; tail -c 300 gpurun_out/r2_b11.err
$B --ms-flags 64 > gpurun_out/r2_b11_spec.json 2>> gpurun_out/r2_b11.err
$B --ms-flags 64 --chunk-len 96 > gpurun_out/r2_b11_spec96.json 2>> gpurun_out/r2_b11.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b11*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']; print(f, 'value', round(d['value']/1e9,1), 'single', round(d['impl_detail']['single_stream']['value']/1e9,1), r['kernel_ms']['ms_fused'], 'frac', round(r['frac'],3), r['events_per_base'])
    except Exception as ex: print(f, 'ERR', ex)
PY
