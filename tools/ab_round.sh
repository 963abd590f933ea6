# A/B of kernel builds on one GPU: kernel-only C4 rates of every library variant, then the parity suites on the product build
set -x
mkdir -p gpurun_out/ab
L=mantaray_b200
LIBS="$L/libmantaray_b200_base.so $L/libmantaray_b200.so $L/libmantaray_b200_nomagic.so $L/libmantaray_b200_nested.so $L/libmantaray_b200_nestednomagic.so $L/libmantaray_b200_b6.so $L/libmantaray_b200_b8.so"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/ab/smi.txt
python tools/kbench.py --rays 1000000 --steps 2048 $LIBS > gpurun_out/ab/kbench_c4_1m.log 2>&1
python tools/kbench.py --rays 1000000 --steps 2048 $L/libmantaray_b200_base.so $L/libmantaray_b200.so >> gpurun_out/ab/kbench_c4_1m.log 2>&1
python tools/kbench.py --workload C3 --rays 1000000 --steps 2048 --notraj $L/libmantaray_b200_base.so $L/libmantaray_b200.so $L/libmantaray_b200_nested.so > gpurun_out/ab/kbench_c3_fin.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_api.py tests/test_gpu_env.py tests/test_golden.py tests/test_reference_behaviour.py -m gpu -x -q --durations=8 > gpurun_out/ab/pytest_subset.log 2>&1
tail -5 gpurun_out/ab/pytest_subset.log
cat gpurun_out/ab/kbench_c4_1m.log gpurun_out/ab/kbench_c3_fin.log
