JH_OPTS=vit_train_fast=0 python profiles/bench_configs.py 1.0 cfg2_viterbi_training_estep_s8_k2 2>&1 | tail -4 | cut -c1-300
