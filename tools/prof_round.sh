set -x
R=${1:-l}
mkdir -p gpurun_out/$R
# launch list of the bench command
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/$R/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/$R/bench_under_ncu.log 2>&1
# full capture of the trace kernel (262144 rays)
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -o gpurun_out/$R/trace -f python bench.py --rays-per-gpu 262144 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/$R/trace.log 2>&1
# DRAM traffic of one 1M-ray launch
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:trace_kernel -s 3 -c 1 --csv --log-file gpurun_out/$R/traffic_1m_rays.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/$R/traffic.log 2>&1
# environment kernel: full capture of the first (all planes) launch, 262144 rays
ncu --set full --clock-control none --import-source on -k regex:sample_kernel -s 1 -c 1 -o gpurun_out/$R/env -f python tools/envbench.py --rays 262144 > gpurun_out/$R/env.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sample_kernel -s 1 -c 1 --csv --log-file gpurun_out/$R/env_traffic_1m.csv python tools/envbench.py > gpurun_out/$R/env_traffic.log 2>&1
# the other workloads, kernel only
for w in C2 C3 C5; do python tools/kbench.py --workload $w --rays 1000000 --steps 2048 >> gpurun_out/$R/workloads.txt 2>&1; done
python tools/kbench.py --workload C3 --rays 1000000 --steps 2048 --notraj >> gpurun_out/$R/workloads.txt 2>&1
ls -la gpurun_out/$R
