set -x
mkdir -p gpurun_out/ab2
L=mantaray_b200
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/tmp_repro.py > gpurun_out/ab2/sanitizer.log 2>&1
grep -v "^=========     Host Frame\|^=========         in " gpurun_out/ab2/sanitizer.log | head -40
LIBS="$L/libmantaray_b200_base.so $L/libmantaray_b200_basemagic.so $L/libmantaray_b200_flatt0.so $L/libmantaray_b200_nestedt0.so $L/libmantaray_b200_flatt0b6.so $L/libmantaray_b200.so"
python tools/kbench.py --rays 1000000 --steps 2048 $LIBS > gpurun_out/ab2/kbench_c4_1m.log 2>&1
cat gpurun_out/ab2/kbench_c4_1m.log
