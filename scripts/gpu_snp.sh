# K1's duration against the mismatch rate of the queries (the workload's is 1 %)
for R in 0 0.001 0.003 0.01 0.03; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batches 4 --snp-rate $R > gpurun_out/snp_$R.json 2> gpurun_out/snp_$R.err; tail -c 200 gpurun_out/snp_$R.err
done
python - <<'PY'
import json
for R in ('0', '0.001', '0.003', '0.01', '0.03'):
    d = json.loads(open('gpurun_out/snp_%s.json' % R).read().strip().split('\n')[-1]); r = d['roofline']
    print('snp', R, 'value', round(d['value']/1e9, 1), 'single', round(d['impl_detail']['single_stream']['value']/1e9, 1), 'K1 us', round(1e3*r['kernel_ms']['ms'], 1), r['events_per_base'])
PY
