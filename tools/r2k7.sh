set -x
R=${1:-r2k7}
mkdir -p gpurun_out/$R
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/$R/pytest_gpu.log 2>&1; tail -3 gpurun_out/$R/pytest_gpu.log
# flags: 0 default, 16 no current map, 8 current map forced
for wf in "C2 0" "C2 16" "C3 0" "C3 16" "C4 0" "C4 8" "C5 0" "C5 8"; do set -- $wf; python tools/kbench.py --workload $1 --rays 1000000 --steps 2048 --flags $2 mantaray_b200/libmantaray_b200.so >> gpurun_out/$R/kbench.log 2>&1; done
cat gpurun_out/$R/kbench.log
