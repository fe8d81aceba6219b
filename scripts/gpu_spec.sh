# K1 with the link words loaded together with the probe (ms flags 64 / 128 / 192: up to depth 14 / 12 / 16) against the default
export PATH=/usr/local/cuda/bin:$PATH
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spec and (find_batch or tiny or matches_batch or randomised)" 2>&1 | tail -2
for F in 0 64 128 192; do
  python bench.py --steps 40 --warmup 5 --no-cpu-baseline --ms-flags $F > gpurun_out/r2j_flags$F.json 2> gpurun_out/r2j_flags$F.err
done
python - <<'PY'
import json
for f in ['r2j_flags0', 'r2j_flags64', 'r2j_flags128', 'r2j_flags192']:
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().split('\n')[-1]); r = d['roofline']
        print(f, 'value', round(d['value']/1e9, 2), 'e2e', round(d['e2e']['value']/1e9, 2), 'K1 us', round(1e3*r['kernel_ms']['ms'], 1), 'frac', round(r['frac'], 3), 'single', round(d['impl_detail']['single_stream']['value']/1e9, 1))
    except Exception as ex:
        print(f, 'ERR', ex)
PY
