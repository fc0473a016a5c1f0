mkdir -p gpurun_out
timeout 900 python scripts/ab2.py scripts/variants/mb8.so scripts/variants/mb12.so > gpurun_out/ab_mb12_r3f.txt 2>&1; tail -4 gpurun_out/ab_mb12_r3f.txt
timeout 900 python scripts/ab_src.py scripts/variants/mb8.so scripts/variants/mb12.so > gpurun_out/ab_src_mb12_r3f.txt 2>&1; tail -4 gpurun_out/ab_src_mb12_r3f.txt | cut -c1-330
