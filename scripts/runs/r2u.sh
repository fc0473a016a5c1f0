mkdir -p gpurun_out
timeout 900 python scripts/ab_src.py scripts/variants/unroll0.so > gpurun_out/ab_src_r2u.txt 2>&1; tail -3 gpurun_out/ab_src_r2u.txt
