# Round 1, late session: parity + gVCF-mode bench after inlining the reference-allele scorer and tabulating the somatic GQ tail.
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r1b_pytest.txt
python bench.py --steps 20 --warmup 3 --gvcf 1 --no-e2e --no-cpu-baseline > gpurun_out/r1b_bench_gvcf.json 2> gpurun_out/r1b_err.txt
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1b_bench_vcf.json 2>> gpurun_out/r1b_err.txt
ncu --set full --clock-control none --import-source on -k regex:pileup_nib -s 4 -c 1 -o gpurun_out/r1b_nib_gvcf -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --gvcf 1 > /dev/null 2>&1
cat gpurun_out/r1b_pytest.txt gpurun_out/r1b_bench_gvcf.json gpurun_out/r1b_bench_vcf.json
