mkdir -p gpurun_out
timeout 900 python scripts/sweep_bg_choice.py > gpurun_out/sweep_bg_choice_r3s.txt 2>&1; tail -6 gpurun_out/sweep_bg_choice_r3s.txt | cut -c1-300
