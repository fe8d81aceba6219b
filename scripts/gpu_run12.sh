python -m pytest tests -m gpu -x -q -k "map_unrefined or golden_map" 2>&1 | tail -3
KBO_BUILD_TIMING=1 python bench.py --config 4 --assemblies 2 > gpurun_out/r2_c4b.json 2> gpurun_out/r2_c4b.err; grep "kbo build" gpurun_out/r2_c4b.err | tail -24
python -c "
import json; d=json.loads(open('gpurun_out/r2_c4b.json').read().strip().split(chr(10))[-1]); print(d['ms_per_step'], d['impl_detail']['split_ms_per_assembly'])"
