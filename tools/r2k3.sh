set -x
R=r2k3
mkdir -p gpurun_out/$R
MANTARAY_B200_LIB=mantaray_b200/libmantaray_b200_s9.so timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/$R/pytest_gpu_s9.log 2>&1; tail -3 gpurun_out/$R/pytest_gpu_s9.log
for L in mantaray_b200/libmantaray_b200_s8.so mantaray_b200/libmantaray_b200_s8a.so mantaray_b200/libmantaray_b200_s9.so; do
  for wf in "C4 4" "C4 6" "C5 6" "C2 4" "C3 4"; do set -- $wf; python tools/kbench.py --workload $1 --rays 1000000 --steps 2048 --flags $2 $L >> gpurun_out/$R/kbench.log 2>&1; done
done
python tools/kbench.py --workload C4 --rays 1000000 --steps 2048 --flags 4 mantaray_b200/libmantaray_b200.so >> gpurun_out/$R/kbench.log 2>&1
cat gpurun_out/$R/kbench.log
