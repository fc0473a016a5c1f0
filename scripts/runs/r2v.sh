mkdir -p gpurun_out
timeout 900 python scripts/sched_prof.py scripts/variants/tprof.so > gpurun_out/sched_prof_r2v.txt 2>&1; tail -12 gpurun_out/sched_prof_r2v.txt
