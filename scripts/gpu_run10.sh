B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
$B --ref-len 2500000 > gpurun_out/r2_b10_half.json 2> gpurun_out/r2_b10.err; tail -c 300 gpurun_out/r2_b10.err
$B --ref-len 2500000 --ms-flags 32 > gpurun_out/r2_b10_half_pairs.json 2>> gpurun_out/r2_b10.err
$B --ref-len 2500000 --ms-flags 16 > gpurun_out/r2_b10_half_fused.json 2>> gpurun_out/r2_b10.err
$B --ref-len 1000000 > gpurun_out/r2_b10_1m.json 2>> gpurun_out/r2_b10.err
$B --ref-len 1000000 --ms-flags 32 > gpurun_out/r2_b10_1m_pairs.json 2>> gpurun_out/r2_b10.err
$B --ref-len 1000000 --ms-flags 16 > gpurun_out/r2_b10_1m_fused.json 2>> gpurun_out/r2_b10.err
$B --ms-flags 32 --no-prefix-table > gpurun_out/r2_b10_pairs_nopref.json 2>> gpurun_out/r2_b10.err
$B --no-prefix-table > gpurun_out/r2_b10_nopref.json 2>> gpurun_out/r2_b10.err
$B --no-rank2 > gpurun_out/r2_b10_norank2.json 2>> gpurun_out/r2_b10.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b10*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']; print(f, 'idx MB', d['impl_detail']['index_device_bytes']>>20, 'value', round(d['value']/1e9,1), 'single', round(d['impl_detail']['single_stream']['value']/1e9,1), 'k1 ms', round(r['kernel_ms']['ms_fused'],4), r['events_per_base'])
    except Exception as ex: print(f, 'ERR', ex)
PY
