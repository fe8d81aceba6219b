# configs 4 / 3 (kbo::map / kbo::call over mutated assemblies, one index per assembly) on the 8 GPUs of one box
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --config 4 --assemblies 32 --asm-reps 2 > gpurun_out/r2g_c4_n8.json 2> gpurun_out/r2g_c4_n8.err; tail -c 300 gpurun_out/r2g_c4_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 8 --config 3 --assemblies 32 --asm-reps 2 > gpurun_out/r2g_c3_n8.json 2> gpurun_out/r2g_c3_n8.err; tail -c 300 gpurun_out/r2g_c3_n8.err
python - <<'PY'
import json
for f in ('r2g_c4_n8', 'r2g_c3_n8'):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().split('\n')[-1]); i = d['impl_detail']
        print(f, 'gpus', d['n_gpus'], 'threads', i['host_threads'], 'ms/asm/rank', round(d['ms_per_step'], 1), 'M bases/s', round(d['value']/1e6), '| one thread', round(i['one_host_thread']['value']/1e6), '| ref index once', round(i['reference_index_built_once']['value']/1e6))
    except Exception as ex:
        print(f, 'ERR', ex)
PY
