# round 2, last run: full GPU tests, smoke, both bench arms at the driver's flags, the default-flag bench, the
# block-staged K1 variant (ms flags 128), serialize/load + K1 variants under compute-sanitizer memcheck / racecheck
set -x
export PATH=/usr/local/cuda/bin:$PATH
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2i_tests.log 2>&1; tail -3 gpurun_out/r2i_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2i_ref.json 2> gpurun_out/r2i_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_1gpu.json 2> gpurun_out/r2i_1gpu.err; tail -c 300 gpurun_out/r2i_1gpu.err
python bench.py > gpurun_out/r2i_default.json 2> gpurun_out/r2i_default.err
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --ms-flags 128 > gpurun_out/r2i_flags128.json 2> gpurun_out/r2i_flags128.err
T="tests/test_gpu_parity.py"
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $T -m gpu -x -q -k "(tiny or serialize or find_batch) and (gated or bstage or serialize)" > gpurun_out/r2i_sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2i_sanitize_memcheck.log
timeout 150 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest $T -m gpu -x -q -k "tiny and (gated or bstage)" > gpurun_out/r2i_sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2i_sanitize_racecheck.log
grep "ERROR SUMMARY" gpurun_out/r2i_sanitize_*.log | sort | uniq -c | head
python - <<'PY'
import json
for f in ['r2i_ref', 'r2i_1gpu', 'r2i_default', 'r2i_flags128']:
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().split('\n')[-1])
        print(f, 'steps', d['steps'], 'value', round(d['value']/1e9, 2), 'e2e', round(d['e2e']['value']/1e9, 2), 'ms/step', round(d['ms_per_step'], 4), 'launches', d.get('gpu_launches'))
        if 'roofline' in d:
            r = d['roofline']; print('   ', r['bound'], 'frac', round(r['frac'], 3), r['kernel_ms'].get('ms'), 'single', round(d['impl_detail']['single_stream']['value']/1e9, 1), d.get('cpu_baseline', {}).get('value'))
    except Exception as ex:
        print(f, 'ERR', ex)
PY
