for i in 1 2; do
python bench.py --config 4 > gpurun_out/p_c4_$i.json 2> gpurun_out/p_c4_$i.err; tail -c 200 gpurun_out/p_c4_$i.err
done
python bench.py --config 3 --k 51 > gpurun_out/p_c3_k51.json 2> gpurun_out/p_c3_k51.err
python - <<'PY'
import json
for f in ('p_c4_1', 'p_c4_2', 'p_c3_k51'):
    d = json.loads(open('gpurun_out/%s.json' % f).read().strip().split('\n')[-1]); i = d['impl_detail']
    print(f, 'ms/asm', round(d['ms_per_step'], 1), '| one thread', round(i['one_host_thread']['ms_per_assembly'], 1), '| ref index once', round(i['reference_index_built_once']['ms_per_assembly'], 1))
PY
nproc; cat /proc/loadavg
