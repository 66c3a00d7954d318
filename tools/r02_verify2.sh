set -x
cd $GRAFT_REPO_ROOT
LPR_PERSIST_PROF=1 timeout 200 python tools/persist_probe.py > gpurun_out/r2_persist_prof4.log 2>&1; tail -4 gpurun_out/r2_persist_prof4.log
timeout 400 python bench.py > gpurun_out/r2_bench_v2.json 2> gpurun_out/r2_bench_v2.err; tail -c 300 gpurun_out/r2_bench_v2.json; tail -3 gpurun_out/r2_bench_v2.err
timeout 700 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 300 > gpurun_out/r2_test_v2.log 2>&1; tail -4 gpurun_out/r2_test_v2.log
