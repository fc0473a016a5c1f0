mkdir -p gpurun_out
timeout 300 python scripts/split_diag.py > gpurun_out/split_diag_r2r.txt 2>&1; grep -c "stats equal True, differing entries 0" gpurun_out/split_diag_r2r.txt; grep "nS" gpurun_out/split_diag_r2r.txt; tail -3 gpurun_out/split_diag_r2r.txt
timeout 300 python scripts/split_prof.py > gpurun_out/split_prof_r2r.txt 2>&1; cat gpurun_out/split_prof_r2r.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "split_kernel or fused or batched_cosmologies or retcodes or lanes" 2>&1 | tail -3
timeout 300 python scripts/ab2.py symboltz.jl_b200/_build/l10_x4_lcdm/libsbm_l10_x4_lcdm.so 2>&1 | tail -2
