# 2-GPU run r3q: bench under torchrun with the final build (headline + strong-scaling extra + sharded config-4 sweep)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --config4 1024 > gpurun_out/bench_n2_r3q.json 2> gpurun_out/bench_n2_r3q.err; tail -c 2600 gpurun_out/bench_n2_r3q.json; tail -3 gpurun_out/bench_n2_r3q.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/dist_capi_check.py > gpurun_out/dist_capi_r3q.txt 2>&1; tail -6 gpurun_out/dist_capi_r3q.txt
