python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -m gpu -x -k "viterbi or band or deterministic or megabase" > gpurun_out/r2Q_pytest.log 2>&1; echo rc=$?; tail -4 gpurun_out/r2Q_pytest.log | cut -c1-250
python profiles/bench_configs.py 1.0 cfg5_viterbi_u8_s128_banded_1Mb full_mid_cfg5_viterbi_u8 full_cfg5_viterbi_u8 2>gpurun_out/r2Q.err | cut -c1-180
JH_OPTS=overlap=0 python profiles/bench_configs.py 1.0 full_mid_cfg5_viterbi_u8 2>/dev/null | cut -c1-180
