python bench.py --config 3 --assemblies 4 --k 51 > gpurun_out/r2_c3_k51.json 2> gpurun_out/r2_c3k.err; tail -c 300 gpurun_out/r2_c3k.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_c3_k51.json').read().strip().split(chr(10))[-1]); print(d['ms_per_step'], d['impl_detail']['split_ms_per_assembly'], d['impl_detail']['parity'])"
