# final single-GPU lines of round 2 (K1 with the LINK_SLOW flag in the link words and the probe-first loop):
# tests, smoke, reference arm, default bench, the K1 variants (ms flags 64: gated contractions, 128: block-staged output,
# 192: both), launch list, ncu --set full of K1
set -x
export PATH=/usr/local/cuda/bin:$PATH
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2h_tests.log 2>&1; tail -3 gpurun_out/r2h_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2h_ref.json 2> gpurun_out/r2h_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_1gpu.json 2> gpurun_out/r2h_1gpu.err; tail -c 300 gpurun_out/r2h_1gpu.err
python bench.py --no-cpu-baseline > gpurun_out/r2h_default.json 2> gpurun_out/r2h_default.err
for F in 64 128 192; do
  python bench.py --steps 40 --warmup 5 --no-cpu-baseline --ms-flags $F > gpurun_out/r2h_flags$F.json 2> gpurun_out/r2h_flags$F.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 4 --warmup 2 --no-cpu-baseline --streams 1 > gpurun_out/r2h_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ms_kernel -s 6 -c 1 -f -o gpurun_out/prof_k1_r2h \
    python bench.py --steps 3 --warmup 3 --streams 1 --no-cpu-baseline > gpurun_out/prof_k1_r2h.log 2>&1
python - <<'PY'
import json
for f in ['r2h_ref', 'r2h_1gpu', 'r2h_default', 'r2h_flags64', 'r2h_flags128', 'r2h_flags192']:
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().split('\n')[-1])
        print(f, 'steps', d['steps'], 'value', round(d['value']/1e9, 2), 'e2e', round(d['e2e']['value']/1e9, 2), 'ms/step', round(d['ms_per_step'], 4), 'launches', d.get('gpu_launches'))
        if 'roofline' in d:
            r = d['roofline']; print('   ', r['bound'], 'frac', round(r['frac'], 3), r['kernel_ms'].get('ms'), 'single', round(d['impl_detail']['single_stream']['value']/1e9, 1), d.get('cpu_baseline', {}).get('value'))
    except Exception as ex:
        print(f, 'ERR', ex)
PY
