# K1 with one lookup per shallow failure: parity at several table depths, then the depth sweep on config 2
set -x
( time python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "prefix or query_sbwt or randomised or golden" ) > gpurun_out/r2c_tests.log 2>&1; tail -4 gpurun_out/r2c_tests.log
for P in 10 11 12 13 0; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --prefix-len $P > gpurun_out/r2c_p$P.json 2> gpurun_out/r2c_p$P.err; tail -c 300 gpurun_out/r2c_p$P.err
done
python - <<'PY'
import json
for P in (10, 11, 12, 13, 0):
    try:
        d = json.loads(open('gpurun_out/r2c_p%d.json' % P).read().strip().split('\n')[-1])
        r = d['roofline']
        print('P', P, 'value', round(d['value']/1e9, 1), 'single', round(d['impl_detail']['single_stream']['value']/1e9, 1), 'e2e', round(d['e2e']['value']/1e9, 1),
              'K1 us', round(1e3*r['kernel_ms']['ms'], 1), 'frac', round(r['frac'], 3), 'B/base', round(r['algorithmic_bytes_per_base'], 1),
              'ref-alg GB/s', round(r['achieved_reference_algorithm_GBps']), 'events', r['events_per_base'], 'bytes', d['impl_detail']['index_device_bytes'])
    except Exception as ex:
        print('P', P, 'ERR', ex)
PY
