python -m pytest tests -q -m gpu --durations=6 > gpurun_out/r2v_pytest.log 2>&1; echo rc=$?; tail -14 gpurun_out/r2v_pytest.log | cut -c1-300
python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; cat gpurun_out/r2v_bench.json | cut -c1-2500; tail -3 gpurun_out/r2v_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2v_bench_ref.json 2> gpurun_out/r2v_bench_ref.err; cat gpurun_out/r2v_bench_ref.json | cut -c1-800
