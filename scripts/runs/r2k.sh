# GPU run r2k: A/B of the slot-major basis sweep and the reciprocal error norm
mkdir -p gpurun_out
python scripts/ab2.py scripts/variants/r2k_base.so scripts/variants/r2k_bslot.so scripts/variants/r2k_normrcp.so scripts/variants/r2k_both.so > gpurun_out/ab_bslot_r2k.txt 2>&1; cat gpurun_out/ab_bslot_r2k.txt
