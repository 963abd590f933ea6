set -x
R=${1:-l}
mkdir -p gpurun_out/$R
python bench.py > gpurun_out/$R/bench_1gpu.json 2> gpurun_out/$R/bench_1gpu.err
cat gpurun_out/$R/bench_1gpu.json
# the same kernel without the depth-floor map (what profiles/r1/m_* describe)
python bench.py --deep-map off --no-e2e --no-cpu > gpurun_out/$R/bench_1gpu_nomap.json 2> gpurun_out/$R/bench_1gpu_nomap.err
cat gpurun_out/$R/bench_1gpu_nomap.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/$R/bench_reference.json 2> gpurun_out/$R/bench_reference.err
cat gpurun_out/$R/bench_reference.json
bash tools/prof_round.sh $R > gpurun_out/$R/prof.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/$R/smoke.log 2>&1; tail -3 gpurun_out/$R/smoke.log
ls -la gpurun_out/$R
