# device refinement (refine.cuh): the whole GPU suite, then configs 3 / 4 (host threads, reference index reuse), default bench
set -x
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2b_tests.log 2>&1; tail -5 gpurun_out/r2b_tests.log
python bench.py --config 4 > gpurun_out/r2b_c4.json 2> gpurun_out/r2b_c4.err; tail -c 300 gpurun_out/r2b_c4.err; head -c 2500 gpurun_out/r2b_c4.json
python bench.py --config 3 > gpurun_out/r2b_c3.json 2> gpurun_out/r2b_c3.err; tail -c 300 gpurun_out/r2b_c3.err; head -c 2500 gpurun_out/r2b_c3.json
python bench.py --config 3 --k 51 > gpurun_out/r2b_c3_k51.json 2> gpurun_out/r2b_c3_k51.err; tail -c 300 gpurun_out/r2b_c3_k51.err; head -c 2500 gpurun_out/r2b_c3_k51.json
