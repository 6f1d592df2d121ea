O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_r02_8gpu.json 2> $O/bench_r02_8gpu.err; tail -2 $O/bench_r02_8gpu.err
