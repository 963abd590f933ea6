# 2-GPU call: the library's own multi-device path (one process, one handle, shared slab queue), executed and measured
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/r2_multi2.sh r2d'
set -x
R=${1:-r2d}
O=gpurun_out/$R
mkdir -p $O
nvidia-smi -L > $O/devices.txt
timeout 600 python -m pytest tests/test_gpu_api.py -q -k "multi_device or pitch" -rs > $O/multi_device_pytest.log 2>&1; tail -4 $O/multi_device_pytest.log
python tools/pciebench.py --gib 4 > $O/pcie_d2h_2gpu.txt 2>&1; tail -8 $O/pcie_d2h_2gpu.txt
MR_DEBUG_TIMING=1 python bench.py --inproc --gpus 1 --total-rays 480000 > $O/inproc_c4_1gpu.json 2> $O/inproc_c4_1gpu.err; cat $O/inproc_c4_1gpu.json
MR_DEBUG_TIMING=1 python bench.py --inproc --gpus 2 --total-rays 480000 > $O/inproc_c4_2gpu.json 2> $O/inproc_c4_2gpu.err; cat $O/inproc_c4_2gpu.json
python bench.py --inproc --gpus 1 --workload C5 --total-rays 8388608 > $O/inproc_c5_1gpu.json 2> $O/inproc_c5_1gpu.err; cat $O/inproc_c5_1gpu.json
python bench.py --inproc --gpus 2 --workload C5 --total-rays 8388608 > $O/inproc_c5_2gpu.json 2> $O/inproc_c5_2gpu.err; cat $O/inproc_c5_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; cat $O/bench_2gpu.json | cut -c1-3000
ls -la $O
