# run-to-run distribution of the hot-kernel time (same binary, same box): default workload vs no explicit candidates
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["ms_per_step"],4), round(d["roofline"]["kernel_ms"],4), d["clocks"])'
for rep in 1 2 3 4 5; do
  python bench.py --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$P" "default"
  python bench.py --no-e2e --no-cpu-baseline --indel-rate 0 2>/dev/null | python -c "$P" "no-indels"
done
