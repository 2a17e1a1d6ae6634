python -m pytest tests -q -m gpu > gpurun_out/r2U_pytest.log 2>&1; echo rc=$?; tail -3 gpurun_out/r2U_pytest.log | cut -c1-250
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2U_smoke.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/r2U_smoke.log
python bench.py > gpurun_out/r2U_bench.json 2> gpurun_out/r2U_bench.err; echo bench rc=$?; cut -c1-400 gpurun_out/r2U_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2U_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2U_ncu.log 2>&1; echo ncu rc=$?
