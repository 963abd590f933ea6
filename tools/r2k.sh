# kernel A/B call: suite, then every (workload, flags) pair on the libraries given
set -x
R=${1:-r2k}; shift
mkdir -p gpurun_out/$R
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/$R/pytest_gpu.log 2>&1; tail -5 gpurun_out/$R/pytest_gpu.log
for L in "$@"; do
for w in C4 C5 C2 C3; do
  for f in 0 4 2 6; do python tools/kbench.py --workload $w --rays 1000000 --steps 2048 --flags $f $L >> gpurun_out/$R/kbench.log 2>&1; done
done
done
cat gpurun_out/$R/kbench.log
