python profiles/bench_configs.py 1.0 cfg5_loglik_s128_banded_1Mb cfg5_bw_s128_banded_1Mb 2>/dev/null | cut -c1-170
