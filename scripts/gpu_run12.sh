python -m pytest tests -m gpu -x -q -k "map or call or config1 or multi_gpu or find_batch" 2>&1 | tail -3
python bench.py --config 4 --assemblies 6 > gpurun_out/r2_c4b.json 2> gpurun_out/r2_c4b.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_c4b.json').read().strip().split(chr(10))[-1]); print(d['ms_per_step'], d['impl_detail']['split_ms_per_assembly'], d['impl_detail']['parity'])"
python bench.py --config 3 --assemblies 6 > gpurun_out/r2_c3b.json 2> gpurun_out/r2_c3b.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_c3b.json').read().strip().split(chr(10))[-1]); print(d['ms_per_step'], d['impl_detail']['split_ms_per_assembly'], d['impl_detail']['parity'])"
KBO_BUILD_TIMING=1 python bench.py --config 4 --assemblies 2 2>&1 >/dev/null | grep "map:\|kbo build" | tail -14
