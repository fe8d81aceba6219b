export PATH=/usr/local/cuda/bin:$PATH
T="tests/test_gpu_parity.py"
K="golden or tiny or find_batch or multi_gpu or many_records or device_pointer"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $T -m gpu -x -q -k "$K" > gpurun_out/r2_sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2_sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest $T -m gpu -x -q -k "golden or tiny or many_records" > gpurun_out/r2_sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2_sanitize_racecheck.log
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest $T -m gpu -x -q -k "golden or tiny or many_records" > gpurun_out/r2_sanitize_initcheck.log 2>&1; echo "initcheck rc=$?"; tail -4 gpurun_out/r2_sanitize_initcheck.log
grep -c "ERROR SUMMARY" gpurun_out/r2_sanitize_*.log; grep "ERROR SUMMARY" gpurun_out/r2_sanitize_*.log | sort | uniq -c | head
