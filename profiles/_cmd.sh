python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "viterbi_training" 2>&1 | tail -1
python profiles/bench_configs.py 1.0 cfg2_viterbi_training_estep_s8_k2 2>/dev/null | cut -c1-200
