python -m pytest tests -q -m gpu > gpurun_out/r3A_pytest.log 2>&1; echo rc=$?; tail -3 gpurun_out/r3A_pytest.log | cut -c1-250
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3A_smoke.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/r3A_smoke.log
python bench.py > gpurun_out/r3A_bench.json 2> gpurun_out/r3A_bench.err; echo bench rc=$?; cut -c1-300 gpurun_out/r3A_bench.json
