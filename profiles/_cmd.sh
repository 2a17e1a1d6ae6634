python -m pytest tests/test_shim_replay.py -q -m gpu > gpurun_out/r2B_pytest.log 2>&1; echo rc=$?; tail -8 gpurun_out/r2B_pytest.log | cut -c1-300
