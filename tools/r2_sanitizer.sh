# compute-sanitizer memcheck over the fast / strict / depth-floor-map / same-grid / uniform-current-map / environment kernels on a subset
# of the parity suites (small cases: the tool slows kernels ~50x)
#   gpurun --timeout 900 -- 'bash tools/r2_sanitizer.sh r2s'
set -x
R=${1:-r2s}
O=gpurun_out/$R
mkdir -p $O
SAN="compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20"
run() { # name, pytest args...
  n=$1; shift
  timeout 600 $SAN python -m pytest "$@" -q -x -p no:cacheprovider > $O/san_$n.log 2>&1
  echo "exit $? : $n" >> $O/sanitizer.log
  grep -E "ERROR SUMMARY|passed|failed" $O/san_$n.log >> $O/sanitizer.log
}
: > $O/sanitizer.log
MR_FUZZ_SEEDS=8 run fuzz tests/test_gpu_fuzz.py
MR_FUZZ_SEEDS=6 run deepmap_fuzz tests/test_gpu_deep_map.py -k "fuzz or steep"
run deepmap tests/test_gpu_deep_map.py -k "dry or default or C4"
run samegrid tests/test_gpu_same_grid.py -k "64-48 or look_alike or C5"
run currentmap tests/test_gpu_current_map.py -k "patchwork or C2 or C5"
run env tests/test_gpu_env.py -k "not full"
run parity tests/test_gpu_parity.py -k "special or infinite or empty or analytic or shoreline"
run api tests/test_gpu_api.py -k "grid_lines or non_affine or negative or pitch or multiple"
cat $O/sanitizer.log
