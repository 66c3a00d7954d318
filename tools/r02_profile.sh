set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 300 > gpurun_out/r2_test8_full.log 2>&1; tail -6 gpurun_out/r2_test8_full.log
timeout 400 python bench.py > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; tail -c 1500 gpurun_out/r2_bench4.json; tail -3 gpurun_out/r2_bench4.err
# ---- ncu: launch lists
LPR_BENCH_SKIP_BB=1 LPR_BENCH_SKIP_REV=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 3 --max-pivots 640 > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ncu_smoke.log 2>&1; tail -2 gpurun_out/ncu_smoke.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_knap -s 40 -c 400 --csv --log-file gpurun_out/r02_launches_knap.csv python tools/knap_probe.py 1 cfg4 > gpurun_out/ncu_knap.log 2>&1; tail -2 gpurun_out/ncu_knap.log
# ---- ncu: full sets
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_persist -c 1 -o gpurun_out/r02_persist python tools/persist_probe.py > gpurun_out/ncu_persist.log 2>&1; tail -2 gpurun_out/ncu_persist.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_knap_plan -s 200 -c 1 -o gpurun_out/r02_knap_plan python tools/knap_probe.py 1 hard > gpurun_out/ncu_knap_plan.log 2>&1; tail -2 gpurun_out/ncu_knap_plan.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_knap_eval -s 200 -c 1 -o gpurun_out/r02_knap_eval python tools/knap_probe.py 1 hard > gpurun_out/ncu_knap_eval.log 2>&1; tail -2 gpurun_out/ncu_knap_eval.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_rank64_update -s 40 -c 1 -o gpurun_out/r02_rank64 python tools/refactor_probe.py 4096 8192 > gpurun_out/ncu_rank64.log 2>&1; tail -2 gpurun_out/ncu_rank64.log
ls -la gpurun_out/*.ncu-rep | tail
