ncu --set full --clock-control none --import-source on -k regex:ms_fused -s 3 -c 1 -o gpurun_out/prof_fused python bench.py --steps 2 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/prof_fused.log 2>&1
tail -3 gpurun_out/prof_fused.log
ls -la gpurun_out/
