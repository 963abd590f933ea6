# First GPU call of the next round: the opt-in parity tests of the depth-floor map (MR_OPT_DEEP_MAP)
# and its kernel-only rates on the four gridded workloads, beside the default path's.  ~2 GPU-minutes.
#   gpurun --timeout 900 -- 'bash tools/validate_deep_map.sh'
set -x
mkdir -p gpurun_out/dmap
MR_TEST_DEEP_MAP=1 timeout 800 python -m pytest tests/test_gpu_deep_map.py -m gpu -q -x > gpurun_out/dmap/pytest.log 2>&1
tail -5 gpurun_out/dmap/pytest.log
for w in C4 C2 C3 C5; do
  for f in 0 1; do python tools/kbench.py --workload $w --rays 1000000 --steps 2048 --flags $f >> gpurun_out/dmap/kbench.log 2>&1; done
done
cat gpurun_out/dmap/kbench.log
