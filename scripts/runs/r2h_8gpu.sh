# 8-GPU run r2h: the driver's SCALE command line at N = 8 (full config-4 sweep, 4096 cosmologies), and the reference arm under torchrun
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8_r2h.json 2> gpurun_out/bench_n8_r2h.err; tail -c 2200 gpurun_out/bench_n8_r2h.json; tail -3 gpurun_out/bench_n8_r2h.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/bench_ref_n8_r2h.json 2> gpurun_out/bench_ref_n8_r2h.err; tail -c 500 gpurun_out/bench_ref_n8_r2h.json
