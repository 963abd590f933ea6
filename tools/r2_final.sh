# Suite, bench lines, ncu evidence of the kernels bench.py times.   gpurun --timeout 1500 -- 'bash tools/r2_final.sh r2f'
set -x
R=${1:-r2f}
O=gpurun_out/$R
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err; cut -c1-1500 $O/bench_1gpu.json; tail -3 $O/bench_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-600 $O/bench_reference.json
NOX="--no-e2e --no-cpu --no-extra --no-parity"
# launch list of the bench command (per-launch times are cold-cache and serialised: the kernel's SHARE of the step is what must agree)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 $NOX > $O/bench_under_ncu.log 2>&1
# full capture of the trace kernels (262144 rays): C4 (depth-floor map), C5 (plain)
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -o $O/trace_c4 -f python bench.py --rays-per-gpu 262144 --steps 1 --warmup 3 $NOX > $O/trace_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -o $O/trace_c5 -f python bench.py --workload C5 --rays-per-gpu 262144 --steps 1 --warmup 3 $NOX > $O/trace_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -o $O/trace_c3 -f python bench.py --workload C3 --rays-per-gpu 262144 --steps 1 --warmup 3 $NOX > $O/trace_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -o $O/trace_c2 -f python bench.py --workload C2 --rays-per-gpu 262144 --steps 1 --warmup 3 $NOX > $O/trace_c2.log 2>&1
# DRAM traffic of one full-size launch
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:trace_kernel -s 3 -c 1 --csv --log-file $O/traffic_c4_1m_rays.csv python bench.py --steps 1 --warmup 3 $NOX > $O/traffic_c4.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -3 $O/smoke.log
ls -la $O
