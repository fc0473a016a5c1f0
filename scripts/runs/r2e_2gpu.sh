# 2-GPU run r2e: bench under torchrun (headline + strong-scaling extras + sharded config-4 sweep), reference arm under torchrun,
# library communicator (C ABI) against torch.distributed
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --config4 1024 > gpurun_out/bench_n2_r2e.json 2> gpurun_out/bench_n2_r2e.err; tail -c 3000 gpurun_out/bench_n2_r2e.json; tail -5 gpurun_out/bench_n2_r2e.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2_r2e.json 2> gpurun_out/bench_ref_n2_r2e.err; tail -c 600 gpurun_out/bench_ref_n2_r2e.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/dist_capi_check.py > gpurun_out/dist_capi_r2e.txt 2>&1; tail -8 gpurun_out/dist_capi_r2e.txt
