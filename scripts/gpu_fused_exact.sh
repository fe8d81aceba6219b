# the one-pass fused kernel (ms flags 24) against the default path
for F in 24; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --ms-flags $F > gpurun_out/r2d_f$F.json 2> gpurun_out/r2d_f$F.err; tail -c 300 gpurun_out/r2d_f$F.err
done
python - <<'PY'
import json
for f in ('r2d_f24',):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().split('\n')[-1]); r = d['roofline']
        print(f, 'value', round(d['value']/1e9, 1), 'single', round(d['impl_detail']['single_stream']['value']/1e9, 1), 'e2e', round(d['e2e']['value']/1e9, 1),
              'K us', round(1e3*r['kernel_ms']['ms'], 1), 'launches', d['gpu_launches'], r['events_per_base'])
    except Exception as ex:
        print(f, 'ERR', ex)
PY
