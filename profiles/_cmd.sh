python -m pytest tests -q -m gpu > gpurun_out/r2I_pytest.log 2>&1; echo rc=$?; tail -4 gpurun_out/r2I_pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2I_smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/r2I_smoke.log
python bench.py --impl reference > gpurun_out/r2I_bench_ref.json 2> gpurun_out/r2I_bench_ref.err; echo ref rc=$?; cut -c1-300 gpurun_out/r2I_bench_ref.json
python bench.py > gpurun_out/r2I_bench.json 2> gpurun_out/r2I_bench.err; echo bench rc=$?; cut -c1-2600 gpurun_out/r2I_bench.json; tail -3 gpurun_out/r2I_bench.err
