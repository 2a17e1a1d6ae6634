python -m pytest tests/test_sunflower.py -q -m gpu > gpurun_out/r2M_pytest.log 2>&1; echo rc=$?; tail -25 gpurun_out/r2M_pytest.log | cut -c1-300
