# device refinement (refine.cuh): parity tests of map / call / index lookups, then configs 3 / 4 with the build / call split
set -x
( time python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "map or call or fill_gaps or index_build or access_kmer or builders or golden" ) 2>&1 | tail -6
KBO_BUILD_TIMING=1 python bench.py --config 4 --assemblies 4 > gpurun_out/r2b_c4.json 2> gpurun_out/r2b_c4.err; tail -c 2500 gpurun_out/r2b_c4.err; cat gpurun_out/r2b_c4.json | head -c 1500
python bench.py --config 3 --assemblies 6 > gpurun_out/r2b_c3.json 2> gpurun_out/r2b_c3.err; tail -c 300 gpurun_out/r2b_c3.err; cat gpurun_out/r2b_c3.json | head -c 1200
