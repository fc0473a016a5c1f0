# GPU run r2d: all GPU tests, A/B of the TMA-staged table rows and the approximate reciprocals against the default build (same box,
# alternating rounds), cost of the lockstep lanes
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/gputest_r2d.log 2>&1; tail -25 gpurun_out/gputest_r2d.log
python scripts/ab2.py scripts/variants/r2_base.so scripts/variants/r2_tma.so scripts/variants/r2_rcp.so scripts/variants/r2_tmarcp.so > gpurun_out/ab_tma_rcp_r2d.txt 2>&1; cat gpurun_out/ab_tma_rcp_r2d.txt
python scripts/lanes_cost.py > gpurun_out/lanes_cost_r2d.txt 2>&1; cat gpurun_out/lanes_cost_r2d.txt
