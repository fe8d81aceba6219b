set -x
KBO_BUILD_TIMING=2 python bench.py --config 3 --assemblies 4 --asm-threads 1 > gpurun_out/p_c3.json 2> gpurun_out/p_c3.err; tail -n 60 gpurun_out/p_c3.err; head -c 1500 gpurun_out/p_c3.json
