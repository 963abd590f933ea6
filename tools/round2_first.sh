# First GPU call of the next round (~12 GPU-minutes):  gpurun --timeout 1500 -- 'bash tools/round2_first.sh r2a'
#  1. the whole GPU suite on the now-default depth-floor-map path
#  2. bench.py (default and --deep-map off), the reference arm, ncu launch list / full capture / traffic of the default kernel
#  3. how the order of the same C4 rays (which rays share a warp / block / wave) moves the kernel time and, via the
#     ncu traffic pass, the cell-record re-reads
set -x
R=${1:-r2a}
mkdir -p gpurun_out/$R
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/$R/pytest_gpu.log 2>&1; tail -3 gpurun_out/$R/pytest_gpu.log
bash tools/final_round.sh $R
for o in start dir tile tile32; do
  for f in 0 2; do python tools/kbench.py --workload C4 --rays 1000000 --steps 2048 --flags $f --order $o >> gpurun_out/$R/order_kbench.log 2>&1; done
done
cat gpurun_out/$R/order_kbench.log
for o in start tile32; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,gpu__time_duration.sum \
      --clock-control none -k regex:trace_kernel -s 1 -c 1 --csv --log-file gpurun_out/$R/order_${o}_traffic.csv \
      python tools/kbench.py --workload C4 --rays 1000000 --steps 2048 --order $o > gpurun_out/$R/order_${o}_ncu.log 2>&1
done
ls -la gpurun_out/$R
