df -h /tmp | tail -1
timeout 1200 python profiles/shards_scale.py 1.0 > gpurun_out/r2N_shards.json 2> gpurun_out/r2N_shards.err; echo rc=$?; cat gpurun_out/r2N_shards.json | cut -c1-1200; tail -3 gpurun_out/r2N_shards.err
