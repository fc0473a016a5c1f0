mkdir -p gpurun_out
timeout 600 python scripts/warp_prof.py > gpurun_out/warp_prof_r3j.txt 2>&1; cat gpurun_out/warp_prof_r3j.txt
