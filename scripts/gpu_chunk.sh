# K1 alone (serial pass of bench.py) and the single-stream step against the chunk length, final K1
export PATH=/usr/local/cuda/bin:$PATH
for C in 64 96 128; do
  python bench.py --steps 40 --warmup 5 --no-cpu-baseline --chunk-len $C > gpurun_out/r2l_chunk$C.json 2> gpurun_out/r2l_chunk$C.err
done
python - <<'PY'
import json
for c in (64, 96, 128):
    try:
        d = json.loads(open('gpurun_out/r2l_chunk%d.json' % c).read().strip().split('\n')[-1]); r = d['roofline']
        print('chunk', c, 'value', round(d['value']/1e9, 2), 'K1 us', round(1e3*r['kernel_ms']['ms'], 1), 'frac', round(r['frac'], 3), 'single', round(d['impl_detail']['single_stream']['value']/1e9, 1), r['events_per_base'])
    except Exception as ex:
        print(c, 'ERR', ex)
PY
