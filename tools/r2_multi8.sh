# 8-GPU call (charged 8x: keep it short): the host's D2H ceiling, C5 sharded 8 ways (block vs interleaved), the
# in-library path on 8 devices
#   gpurun --gpus 8 --timeout 600 -- 'bash tools/r2_multi8.sh r2e'
set -x
R=${1:-r2e}
O=gpurun_out/$R
mkdir -p $O
python tools/pciebench.py --gib 2 --reps 2 > $O/pcie_d2h_8gpu.txt 2>&1; tail -12 $O/pcie_d2h_8gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518"
# the driver's own command at N = 8 (headline C4 + every named shape + e2e)
$TR bench.py --gpus 8 --steps 3 --warmup 3 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; cut -c1-1200 $O/bench_8gpu.json; tail -2 $O/bench_8gpu.err
$TR bench.py --gpus 8 --workload C5 --steps 2 --warmup 3 --no-extra --shard block --no-e2e > $O/bench_c5_8gpu_block.json 2> $O/bench_c5_8gpu_block.err; cat $O/bench_c5_8gpu_block.json | cut -c1-2500
$TR bench.py --gpus 8 --workload C5 --steps 2 --warmup 3 --no-extra --shard interleave > $O/bench_c5_8gpu_interleave.json 2> $O/bench_c5_8gpu_interleave.err; cat $O/bench_c5_8gpu_interleave.json | cut -c1-2500
MR_DEBUG_TIMING=1 python bench.py --inproc --gpus 8 --workload C5 --total-rays 16777216 --steps 2 > $O/inproc_c5_8gpu.json 2> $O/inproc_c5_8gpu.err; cat $O/inproc_c5_8gpu.json
MR_DEBUG_TIMING=1 python bench.py --inproc --gpus 8 --total-rays 960000 --steps 2 > $O/inproc_c4_8gpu.json 2> $O/inproc_c4_8gpu.err; cat $O/inproc_c4_8gpu.json
timeout 300 python -m pytest tests/test_gpu_api.py -q -k "multi_device" -rs > $O/multi_device_pytest_8gpu.log 2>&1; tail -3 $O/multi_device_pytest_8gpu.log
ls -la $O
