# 8-GPU run r3q: the driver's SCALE command line at N = 8 with the final build (full config-4 sweep, 4096 cosmologies), and the reference arm under torchrun
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8_r3q.json 2> gpurun_out/bench_n8_r3q.err; tail -c 1800 gpurun_out/bench_n8_r3q.json; tail -3 gpurun_out/bench_n8_r3q.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/bench_ref_n8_r3q.json 2> gpurun_out/bench_ref_n8_r3q.err; tail -c 500 gpurun_out/bench_ref_n8_r3q.json
