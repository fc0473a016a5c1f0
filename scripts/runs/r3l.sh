mkdir -p gpurun_out
for v in wprof wprof_nofence; do timeout 600 python scripts/warp_prof.py scripts/variants/$v.so 2>&1 | head -4 | cut -c1-150; done
