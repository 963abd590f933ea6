# second GPU call of round 2: the suite on the new kernels (same-grid shortcut, fourth-root deep path), then A/B
set -x
R=${1:-r2b}
mkdir -p gpurun_out/$R
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/$R/pytest_gpu.log 2>&1; tail -5 gpurun_out/$R/pytest_gpu.log
L=mantaray_b200/libmantaray_b200.so
for w in C4 C5 C2 C3; do
  for f in 0 4 2 6; do python tools/kbench.py --workload $w --rays 1000000 --steps 2048 --flags $f $L >> gpurun_out/$R/kbench.log 2>&1; done
  python tools/kbench.py --workload $w --rays 1000000 --steps 2048 --flags 0 mantaray_b200/libmantaray_b200_b6.so >> gpurun_out/$R/kbench.log 2>&1
done
cat gpurun_out/$R/kbench.log
