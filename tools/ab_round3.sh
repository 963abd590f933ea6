set -x
mkdir -p gpurun_out/ab3
L=mantaray_b200
for m in 0 1; do timeout 300 compute-sanitizer --tool memcheck --print-limit 3 python tools/tmp_repro.py $m > gpurun_out/ab3/sanitizer_$m.log 2>&1; grep -v "Host Frame\|^=========         in " gpurun_out/ab3/sanitizer_$m.log | head -30; done
timeout 900 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/ab3/pytest_all.log 2>&1
tail -15 gpurun_out/ab3/pytest_all.log
python tools/kbench.py --rays 1000000 --steps 2048 $L/libmantaray_b200_base.so $L/libmantaray_b200.so $L/libmantaray_b200_nomagic.so > gpurun_out/ab3/kbench_c4_1m.log 2>&1
cat gpurun_out/ab3/kbench_c4_1m.log
