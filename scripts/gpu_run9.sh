B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
$B --ms-flags 32 > gpurun_out/r2_b9_pairs.json 2>> gpurun_out/r2_b9.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b9_pairs.json')):
    d=json.load(open(f)); r=d['roofline']; print(f, 'value', round(d['value']/1e9,1), 'single', round(d['impl_detail']['single_stream']['value']/1e9,1), r['kernel_ms']['ms_fused'], 'frac', round(r['frac'],3), r['events_per_base'])
PY
ncu --set full --clock-control none --import-source on -k regex:ms_pairs -s 3 -c 1 -o gpurun_out/prof_pairs python bench.py --steps 2 --warmup 1 --no-cpu-baseline --streams 1 --ms-flags 32 > gpurun_out/prof_pairs.log 2>&1
