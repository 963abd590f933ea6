set -x
R=${1:-r2k8}
mkdir -p gpurun_out/$R
L=mantaray_b200/libmantaray_b200.so
for w in C4 C5; do
  python tools/kbench.py --workload $w --rays 1000000 --steps 2048 $L >> gpurun_out/$R/kbench.log 2>&1
  for p in 0.5 1.0; do for m in normal streaming; do
    echo "# MR_L2_PERSIST=$p MR_L2_MISS=$m" >> gpurun_out/$R/kbench.log
    MR_L2_PERSIST=$p MR_L2_MISS=$m python tools/kbench.py --workload $w --rays 1000000 --steps 2048 $L >> gpurun_out/$R/kbench.log 2>&1
  done; done
done
cat gpurun_out/$R/kbench.log
