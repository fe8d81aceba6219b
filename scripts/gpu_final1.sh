# final single-GPU lines of the round: tests, smoke, default bench + reference arm, launch list, fused variants, config 5
set -x
export PATH=/usr/local/cuda/bin:$PATH
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2f_tests.log 2>&1; tail -3 gpurun_out/r2f_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_1gpu.json 2> gpurun_out/r2f_1gpu.err; tail -c 300 gpurun_out/r2f_1gpu.err
python bench.py > gpurun_out/r2f_default.json 2> gpurun_out/r2f_default.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --ms-flags 24 > gpurun_out/r2f_fused_onepass.json 2> gpurun_out/r2f_fused_onepass.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --ms-flags 16 > gpurun_out/r2f_fused_twopass.json 2> gpurun_out/r2f_fused_twopass.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 4 --warmup 2 --no-cpu-baseline --streams 1 > gpurun_out/r2f_launches.log 2>&1
python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/r2f_c5_1g.json 2> gpurun_out/r2f_c5_1g.err; tail -c 300 gpurun_out/r2f_c5_1g.err
python - <<'PY'
import json
for f in ['r2f_ref', 'r2f_1gpu', 'r2f_default', 'r2f_fused_onepass', 'r2f_fused_twopass', 'r2f_c5_1g']:
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().split('\n')[-1])
        print(f, 'steps', d['steps'], 'value', round(d['value']/1e9, 2), 'e2e', round(d['e2e']['value']/1e9, 2), 'ms/step', round(d['ms_per_step'], 4), 'launches', d.get('gpu_launches'))
        if 'roofline' in d:
            r = d['roofline']; print('   ', r['bound'], 'frac', round(r['frac'], 3), r['kernel_ms'].get('ms'), 'single', round(d['impl_detail']['single_stream']['value']/1e9, 1), d.get('cpu_baseline', {}).get('value'))
    except Exception as ex:
        print(f, 'ERR', ex)
PY
