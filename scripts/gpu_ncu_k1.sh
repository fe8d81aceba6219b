# ncu --set full on one launch of K1 (bench.py config 2, one stream) at a given depth of the prefix-state table
export PATH=/usr/local/cuda/bin:$PATH
for P in ${DEPTHS:-11}; do
  ncu --set full --clock-control none --import-source on -k regex:ms_kernel -s 6 -c 1 -f -o gpurun_out/prof_k1_p$P \
      python bench.py --steps 3 --warmup 3 --streams 1 --no-cpu-baseline --prefix-len $P > gpurun_out/prof_k1_p$P.log 2>&1
  tail -2 gpurun_out/prof_k1_p$P.log | cut -c1-200
done
