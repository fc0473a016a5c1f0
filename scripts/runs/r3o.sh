mkdir -p gpurun_out
timeout 900 python scripts/ab2.py scripts/variants/before.so scripts/variants/bf2.so > gpurun_out/ab_bf2_r3o.txt 2>&1; tail -4 gpurun_out/ab_bf2_r3o.txt
timeout 900 python scripts/ab_src.py scripts/variants/before.so scripts/variants/bf2.so > gpurun_out/ab_src_bf2_r3o.txt 2>&1; tail -2 gpurun_out/ab_src_bf2_r3o.txt | cut -c1-330
