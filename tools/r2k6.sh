set -x
R=${1:-r2k6}
mkdir -p gpurun_out/$R
MANTARAY_B200_LIB=mantaray_b200/libmantaray_b200_k0s.so timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/$R/pytest_gpu_k0s.log 2>&1; tail -3 gpurun_out/$R/pytest_gpu_k0s.log
for L in mantaray_b200/libmantaray_b200_k0s.so mantaray_b200/libmantaray_b200_shadow.so mantaray_b200/libmantaray_b200.so; do
for wf in "C4 0" "C5 0" "C3 0" "C2 0" "C4 2"; do set -- $wf; python tools/kbench.py --workload $1 --rays 1000000 --steps 2048 --flags $2 $L >> gpurun_out/$R/kbench.log 2>&1; done
done
cat gpurun_out/$R/kbench.log
