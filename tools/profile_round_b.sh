# Round 1, late session ("r1b"): full GPU test pass, bench lines, launch list, full ncu captures of the hot kernel in both modes.
# The .ncu-rep files are summarised ON the box (profiles/ncu_summary.py) and removed: two full reports exceed the 64 MiB that come back.
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r1b_pytest.txt
python bench.py > gpurun_out/r1b_bench_default.json 2> gpurun_out/r1b_err.txt
python bench.py --gvcf 1 --no-e2e --no-cpu-baseline > gpurun_out/r1b_bench_gvcf.json 2>> gpurun_out/r1b_err.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1b_bench_reference.json 2>> gpurun_out/r1b_err.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pileup|score_|tile_|nib_|unpack_|apply_|reads_|gather_|prune_|DeviceScan|gq_tail" -c 400 --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:pileup_nib -s 4 -c 1 -o /tmp/r1b_nib -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python profiles/ncu_summary.py /tmp/r1b_nib.ncu-rep --sass 25 > gpurun_out/r1b_nib_kernel_ncu.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:pileup_nib -s 4 -c 1 -o /tmp/r1b_nib_gvcf -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --gvcf 1 > /dev/null 2>&1
python profiles/ncu_summary.py /tmp/r1b_nib_gvcf.ncu-rep --sass 25 > gpurun_out/r1b_nib_gvcf_kernel_ncu.txt 2>&1
cat gpurun_out/r1b_pytest.txt; ls -la gpurun_out
