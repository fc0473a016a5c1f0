mkdir -p gpurun_out
timeout 900 python scripts/ab2.py scripts/variants/base.so scripts/variants/bf.so scripts/variants/unroll.so > gpurun_out/ab_stage_r2s.txt 2>&1; tail -7 gpurun_out/ab_stage_r2s.txt
