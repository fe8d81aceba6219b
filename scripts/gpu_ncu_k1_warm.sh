# ncu --set full on one launch of the final K1 with the caches left warm (--cache-control none)
export PATH=/usr/local/cuda/bin:$PATH
ncu --set full --clock-control none --cache-control none --import-source on -k regex:ms_kernel -s 6 -c 1 -f -o gpurun_out/prof_k1_r2m_warm \
    python bench.py --steps 3 --warmup 3 --streams 1 --no-cpu-baseline > gpurun_out/prof_k1_r2m_warm.log 2>&1
tail -1 gpurun_out/prof_k1_r2m_warm.log | cut -c1-200
