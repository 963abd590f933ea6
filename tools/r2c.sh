# third GPU call of round 2: suite on the reworked kernels, then A/B of the knobs
set -x
R=${1:-r2c}
mkdir -p gpurun_out/$R
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/$R/pytest_gpu.log 2>&1; tail -5 gpurun_out/$R/pytest_gpu.log
L=mantaray_b200/libmantaray_b200.so
for w in C4 C5 C2 C3; do
  for f in 0 4 2 6; do python tools/kbench.py --workload $w --rays 1000000 --steps 2048 --flags $f $L >> gpurun_out/$R/kbench.log 2>&1; done
done
for w in C4 C3; do
  for f in 0 4; do python tools/kbench.py --workload $w --rays 1000000 --steps 2048 --flags $f mantaray_b200/libmantaray_b200_nr4.so >> gpurun_out/$R/kbench.log 2>&1; done
done
for f in 4 6; do python tools/kbench.py --workload C4 --rays 1000000 --steps 2048 --flags $f mantaray_b200/libmantaray_b200_r1like.so >> gpurun_out/$R/kbench.log 2>&1; done
cat gpurun_out/$R/kbench.log
