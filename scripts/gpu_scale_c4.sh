# config 4 (kbo::map over mutated assemblies, one index per assembly) on 8 GPUs of one box: waits that sleep, 2 / 4 host threads per rank
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --config 4 --assemblies 16 --blocking-sync 1 > gpurun_out/r2f_c4_n8_block.json 2> gpurun_out/r2f_c4_n8_block.err; tail -c 300 gpurun_out/r2f_c4_n8_block.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 8 --config 4 --assemblies 16 --asm-threads 2 --asm-reps 2 > gpurun_out/r2f_c4_n8_t2.json 2> gpurun_out/r2f_c4_n8_t2.err; tail -c 300 gpurun_out/r2f_c4_n8_t2.err
python - <<'PY'
import json
for f in ('r2f_c4_n8_block', 'r2f_c4_n8_t2'):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().split('\n')[-1]); i = d['impl_detail']
        print(f, 'gpus', d['n_gpus'], 'threads', i['host_threads'], i['host_waits'], 'ms/asm/rank', round(d['ms_per_step'], 1), 'M bases/s', round(d['value']/1e6), '| one thread', round(i['one_host_thread']['value']/1e6), '| ref index once', round(i['reference_index_built_once']['value']/1e6))
    except Exception as ex:
        print(f, 'ERR', ex)
PY
