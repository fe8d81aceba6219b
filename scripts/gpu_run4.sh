B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
$B > gpurun_out/r2_b4.json 2> gpurun_out/r2_b4.err; tail -c 800 gpurun_out/r2_b4.err
$B --chunk-len 96 > gpurun_out/r2_b4_c96.json 2>> gpurun_out/r2_b4.err
$B --chunk-len 40 > gpurun_out/r2_b4_c40.json 2>> gpurun_out/r2_b4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b4*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']; print(f, 'value', round(d['value']/1e9,1), 'single', round(d['impl_detail']['single_stream']['value']/1e9,1), 'e2e', round(d['e2e']['value']/1e9,1), r['kernel_ms']['pack'], r['kernel_ms']['ms_fused'], r['events_per_base'], round(r['algorithmic_bytes_per_base'],1), 'frac', round(r['frac'],3), r['bound'])
    except Exception as ex: print(f, 'ERR', ex)
PY
ncu --set full --clock-control none --import-source on -k regex:ms_fused -s 3 -c 1 -o gpurun_out/prof_fused2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/prof_fused2.log 2>&1
