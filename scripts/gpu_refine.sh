# device refinement (refine.cuh): the whole GPU suite, then configs 3 / 4 (host threads, reference index reuse)
set -x
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2b_tests.log 2>&1; tail -5 gpurun_out/r2b_tests.log
python bench.py --config 4 > gpurun_out/r2b_c4.json 2> gpurun_out/r2b_c4.err; tail -c 300 gpurun_out/r2b_c4.err
python bench.py --config 3 > gpurun_out/r2b_c3.json 2> gpurun_out/r2b_c3.err; tail -c 300 gpurun_out/r2b_c3.err
python bench.py --config 3 --k 51 > gpurun_out/r2b_c3_k51.json 2> gpurun_out/r2b_c3_k51.err; tail -c 300 gpurun_out/r2b_c3_k51.err
python bench.py --config 4 --k 51 > gpurun_out/r2b_c4_k51.json 2> gpurun_out/r2b_c4_k51.err; tail -c 300 gpurun_out/r2b_c4_k51.err
python - <<'PY'
import json
for f in ('r2b_c4', 'r2b_c3', 'r2b_c3_k51', 'r2b_c4_k51'):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().split('\n')[-1]); i = d['impl_detail']
        print(f, 'ms/asm', round(d['ms_per_step'], 1), 'M bases/s', round(d['value']/1e6), '| one thread', round(i['one_host_thread']['ms_per_assembly'], 1),
              {k[:24]: round(v, 1) for k, v in i['one_host_thread']['split_ms_per_assembly'].items()}, '| ref index once', round(i['reference_index_built_once']['ms_per_assembly'], 1), i['parity']['checked'])
    except Exception as ex:
        print(f, 'ERR', ex)
PY
