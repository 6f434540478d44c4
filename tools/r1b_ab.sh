# A/B of hot-kernel variants selected by --tune-prefetch, plus a quick parity pass
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/ab_pytest.txt
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])'
for rep in 1 2; do
for v in 0 8; do python bench.py --steps 30 --warmup 5 --gvcf 1 --no-e2e --no-cpu-baseline --tune-prefetch $v | python -c "$P" "gvcf variant $v"; done
for v in 0 8; do python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --tune-prefetch $v | python -c "$P" "vcf variant $v"; done
done > gpurun_out/ab.txt 2>/dev/null
cat gpurun_out/ab_pytest.txt gpurun_out/ab.txt
