# ncu --set full on one launch of each small kernel of a find step (bench.py config 2, one stream)
export PATH=/usr/local/cuda/bin:$PATH
ncu --set full --clock-control none --import-source on -k regex:'pack_queries|derand_translate_bits|rle_word_counts|rle_finish' -s 12 -c 4 -f -o gpurun_out/prof_small_r2k \
    python bench.py --steps 3 --warmup 3 --streams 1 --no-cpu-baseline > gpurun_out/prof_small_r2k.log 2>&1
tail -2 gpurun_out/prof_small_r2k.log | cut -c1-300
