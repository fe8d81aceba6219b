nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -m pytest tests -m gpu -x -q -k "multi_gpu" 2>&1 | tail -4
for N in 8 4 2 1; do
  if [ $N -gt 1 ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_scale_$N.json 2> gpurun_out/r2_scale_$N.err
  else
    python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_scale_$N.json 2> gpurun_out/r2_scale_$N.err
  fi
  tail -c 300 gpurun_out/r2_scale_$N.err
done
python - <<'PY'
import json
base=None
for N in (1,2,4,8):
    try:
        d=json.loads(open('gpurun_out/r2_scale_%d.json'%N).read().strip().split('\n')[-1])
        if N==1: base=d
        print(N, 'value', round(d['value']/1e9,1), 'eff', round(d['value']/base['value']/N,3), 'e2e', round(d['e2e']['value']/1e9,1), 'eff', round(d['e2e']['value']/base['e2e']['value']/N,3), 'single', round(d['impl_detail']['single_stream']['value']/1e9,1), d['impl_detail'].get('host_affinity'))
    except Exception as ex: print(N,'ERR',ex)
PY
