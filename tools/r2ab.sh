# kernel A/B: C4, C5, C2, C3 at default flags on each library given (twice)
R=${1:-r2ab}; shift
mkdir -p gpurun_out/$R
for rep in 1 2; do
for L in "$@"; do
for w in C4 C5 C2 C3; do
  python tools/kbench.py --workload $w --rays 1000000 --steps 2048 --flags 0 $L >> gpurun_out/$R/kbench.log 2>&1
done
done
done
grep -o '"lib": "[a-z0-9_.]*"\|"workload": "C."\|"ms": [0-9.]*' gpurun_out/$R/kbench.log | paste - - -
