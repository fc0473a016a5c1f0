# GPU run r2m: A/B of the ELL slot permutation (generator only: bank-aware placement of a row's entries in its round's slots)
mkdir -p gpurun_out
python scripts/ab2.py scripts/variants/r2m_base.so scripts/variants/r2m_ellperm.so > gpurun_out/ab_ellperm_r2m.txt 2>&1; cat gpurun_out/ab_ellperm_r2m.txt
