#!/bin/bash
# compute-sanitizer over the GPU test suite (run under gpurun): memcheck on everything, racecheck on the suites that drive the shared-memory kernels
# (hot kernel queue + scorer, amplicon tallies, transpose, gather). Logs -> gpurun_out/r2/ (copied to profiles/r2/).
mkdir -p gpurun_out/r2
O=gpurun_out/r2
compute-sanitizer --tool memcheck --log-file $O/r2b_sanitizer_memcheck.log --print-limit 20 \
    python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/r2b_sanitizer_memcheck_pytest.txt 2>&1
tail -3 $O/r2b_sanitizer_memcheck_pytest.txt; grep -c "ERROR SUMMARY" $O/r2b_sanitizer_memcheck.log; grep "ERROR SUMMARY" $O/r2b_sanitizer_memcheck.log | sort | uniq -c | head
compute-sanitizer --tool racecheck --log-file $O/r2b_sanitizer_racecheck.log --print-limit 20 \
    python -m pytest tests/test_gpu_amplicon.py tests/test_gpu_reads_path.py tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider > $O/r2b_sanitizer_racecheck_pytest.txt 2>&1
tail -3 $O/r2b_sanitizer_racecheck_pytest.txt; grep "RACECHECK SUMMARY" $O/r2b_sanitizer_racecheck.log | sort | uniq -c | head
