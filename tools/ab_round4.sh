set -x
mkdir -p gpurun_out/ab4
L=mantaray_b200
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/ab4/pytest_all.log 2>&1
tail -15 gpurun_out/ab4/pytest_all.log
python tools/kbench.py --rays 1000000 --steps 2048 $L/libmantaray_b200_base.so $L/libmantaray_b200.so $L/libmantaray_b200_ld1.so $L/libmantaray_b200_ld2.so $L/libmantaray_b200_b6.so $L/libmantaray_b200_b8.so > gpurun_out/ab4/kbench_c4_1m.log 2>&1
python tools/kbench.py --rays 927369 --steps 2048 $L/libmantaray_b200.so >> gpurun_out/ab4/kbench_c4_1m.log 2>&1
python tools/kbench.py --rays 1058841 --steps 2048 $L/libmantaray_b200.so >> gpurun_out/ab4/kbench_c4_1m.log 2>&1
cat gpurun_out/ab4/kbench_c4_1m.log
