mkdir -p gpurun_out
timeout 600 python scripts/warp_prof.py scripts/variants/wprof.so > gpurun_out/warp_prof_r3l.txt 2>&1; head -16 gpurun_out/warp_prof_r3l.txt | cut -c1-170
