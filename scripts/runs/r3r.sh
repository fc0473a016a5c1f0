mkdir -p gpurun_out
timeout 900 python scripts/ab2.py scripts/variants/before.so scripts/variants/pre.so > gpurun_out/ab_pre_r3r.txt 2>&1; tail -4 gpurun_out/ab_pre_r3r.txt
timeout 900 python scripts/ab_src.py scripts/variants/before.so scripts/variants/pre.so > gpurun_out/ab_src_pre_r3r.txt 2>&1; tail -2 gpurun_out/ab_src_pre_r3r.txt | cut -c1-330
