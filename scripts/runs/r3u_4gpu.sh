# 4-GPU run r3u: the driver's SCALE command line at N = 4 with the last build
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_n4_r3u.json 2> gpurun_out/bench_n4_r3u.err; tail -c 600 gpurun_out/bench_n4_r3u.json; tail -2 gpurun_out/bench_n4_r3u.err
