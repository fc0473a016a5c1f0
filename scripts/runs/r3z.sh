mkdir -p gpurun_out
timeout 900 python scripts/save_cost.py > gpurun_out/save_cost_r3z.txt 2>&1; cat gpurun_out/save_cost_r3z.txt | cut -c1-200
