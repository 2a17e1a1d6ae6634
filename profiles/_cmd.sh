python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "band" > gpurun_out/r2X_pytest.log 2>&1; echo rc=$?; tail -3 gpurun_out/r2X_pytest.log | cut -c1-250
python profiles/bench_configs.py 1.0 cfg5_loglik_s128_banded_1Mb cfg5_bw_s128_banded_1Mb cfg5_decode_s128_banded_1Mb 2>/dev/null | cut -c1-170
JH_OPTS=band_meet=0 python profiles/bench_configs.py 1.0 cfg5_loglik_s128_banded_1Mb 2>/dev/null | cut -c1-170
