# is K1 slowed by L2 misses?  (a) one batch re-used vs 16 batches rotated; (b) ncu with the caches left warm
export PATH=/usr/local/cuda/bin:$PATH
for B in 1 16; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batches $B > gpurun_out/l2p_b$B.json 2> gpurun_out/l2p_b$B.err; tail -c 200 gpurun_out/l2p_b$B.err
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-l2-persist > gpurun_out/l2p_nopersist.json 2> gpurun_out/l2p_nopersist.err
python - <<'PY'
import json
for f in ('l2p_b1', 'l2p_b16', 'l2p_nopersist'):
    d = json.loads(open('gpurun_out/%s.json' % f).read().strip().split('\n')[-1]); r = d['roofline']
    print(f, 'value', round(d['value']/1e9, 1), 'single', round(d['impl_detail']['single_stream']['value']/1e9, 1), 'K1 us', round(1e3*r['kernel_ms']['ms'], 1), 'pack', round(1e3*r['kernel_ms']['pack'], 1))
PY
ncu --set full --cache-control none --clock-control none --import-source on -k regex:ms_kernel -s 6 -c 1 -f -o gpurun_out/prof_k1_warm \
    python bench.py --steps 3 --warmup 3 --streams 1 --no-cpu-baseline > gpurun_out/prof_k1_warm.log 2>&1
tail -1 gpurun_out/prof_k1_warm.log | cut -c1-100
