set -x
cd $GRAFT_REPO_ROOT
timeout 700 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 300 > gpurun_out/r2_test_final.log 2>&1; tail -6 gpurun_out/r2_test_final.log
timeout 400 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -c 1200 gpurun_out/r2_bench_final.json; tail -3 gpurun_out/r2_bench_final.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2>&1; tail -c 600 gpurun_out/r2_bench_ref.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_knap -s 60 -c 480 --csv --log-file gpurun_out/r02_launches_knap.csv python tools/knap_probe.py 1 cfg4 > gpurun_out/ncu_knap.log 2>&1; tail -2 gpurun_out/ncu_knap.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_knap -s 3000 -c 240 --csv --log-file gpurun_out/r02_launches_knap_hard.csv python tools/knap_probe.py 1 hard > gpurun_out/ncu_knap2.log 2>&1; tail -2 gpurun_out/ncu_knap2.log
