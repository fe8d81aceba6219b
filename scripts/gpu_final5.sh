# round 2, last run on the final tree (packed block scan in rle_word_counts): full GPU tests, smoke, both bench arms at the driver's flags, the default-flag
# bench and the launch list (K2b at 12 blocks per SM, single-thread fence in the RLE counting kernels)
set -x
export PATH=/usr/local/cuda/bin:$PATH
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2m_tests.log 2>&1; tail -3 gpurun_out/r2m_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2m_ref.json 2> gpurun_out/r2m_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_1gpu.json 2> gpurun_out/r2m_1gpu.err; tail -c 300 gpurun_out/r2m_1gpu.err
python bench.py > gpurun_out/r2m_default.json 2> gpurun_out/r2m_default.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/r2m_launches.csv python bench.py --steps 4 --warmup 2 --no-cpu-baseline --streams 1 > gpurun_out/r2m_launches.log 2>&1
python - <<'PY'
import json
for f in ['r2m_ref', 'r2m_1gpu', 'r2m_default']:
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().split('\n')[-1])
        print(f, 'steps', d['steps'], 'value', round(d['value']/1e9, 2), 'e2e', round(d['e2e']['value']/1e9, 2), 'ms/step', round(d['ms_per_step'], 4), 'launches', d.get('gpu_launches'))
        if 'roofline' in d:
            r = d['roofline']; print('   ', r['bound'], 'frac', round(r['frac'], 3), r['kernel_ms'].get('ms'), 'single', round(d['impl_detail']['single_stream']['value']/1e9, 1), d.get('cpu_baseline', {}).get('value'))
    except Exception as ex:
        print(f, 'ERR', ex)
PY
