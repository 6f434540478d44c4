# A/B of hot-kernel variants selected by --tune-prefetch (gVCF mode), plus a quick parity pass of the gVCF tests
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/ab_pytest.txt
for v in 0 8 0 8; do python bench.py --steps 30 --warmup 5 --gvcf 1 --no-e2e --no-cpu-baseline --tune-prefetch $v | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant $v', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"; done > gpurun_out/ab.txt 2>&1
python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('vcf', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])" >> gpurun_out/ab.txt 2>&1
cat gpurun_out/ab_pytest.txt gpurun_out/ab.txt
