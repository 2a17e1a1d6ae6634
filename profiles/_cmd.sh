python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "band or viterbi or long" > gpurun_out/r2w_pytest.log 2>&1; echo rc=$?; tail -3 gpurun_out/r2w_pytest.log | cut -c1-300
python profiles/bench_configs.py 1.0 cfg5_viterbi_u8_s128_banded_1Mb full_mid_cfg5_viterbi_u8 > gpurun_out/r2w_cfg.jsonl 2>gpurun_out/r2w_cfg.err; cut -c1-200 gpurun_out/r2w_cfg.jsonl; tail -2 gpurun_out/r2w_cfg.err
