set -x
cd $GRAFT_REPO_ROOT
timeout 700 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 300 > gpurun_out/r2_test_verify.log 2>&1; tail -4 gpurun_out/r2_test_verify.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke_verify.log 2>&1; tail -2 gpurun_out/r2_smoke_verify.log
timeout 400 python bench.py > gpurun_out/r2_bench_verify.json 2> gpurun_out/r2_bench_verify.err; tail -c 600 gpurun_out/r2_bench_verify.json; tail -3 gpurun_out/r2_bench_verify.err
