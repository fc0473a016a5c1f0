mkdir -p gpurun_out
timeout 900 python scripts/ab2.py scripts/variants/switch0.so scripts/variants/switch1.so > gpurun_out/ab_switch_r3h.txt 2>&1; tail -4 gpurun_out/ab_switch_r3h.txt
timeout 900 python scripts/ab_src.py scripts/variants/switch0.so scripts/variants/switch1.so > gpurun_out/ab_src_switch_r3h.txt 2>&1; tail -4 gpurun_out/ab_src_switch_r3h.txt | cut -c1-330
