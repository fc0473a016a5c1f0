mkdir -p gpurun_out
timeout 900 python scripts/split_stress.py > gpurun_out/split_stress_r3y.txt 2>&1; tail -8 gpurun_out/split_stress_r3y.txt | cut -c1-300
