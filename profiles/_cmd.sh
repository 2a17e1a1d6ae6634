python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "deterministic" > gpurun_out/r2P_pytest.log 2>&1; echo rc=$?; tail -25 gpurun_out/r2P_pytest.log | cut -c1-250
JH_OPTS=deterministic=1 python profiles/bench_configs.py 1.0 cfg2_bw_s8_k2 cfg3_bw_s32_k3 cfg4_bw_s16_k1_ragged 2>&1 | cut -c1-180
python profiles/bench_configs.py 1.0 cfg3_bw_s32_k3 2>&1 | cut -c1-180
