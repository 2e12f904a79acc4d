N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_r02g_n$N.log 2> gpurun_out/bench_r02g_n$N.err; tail -c 300 gpurun_out/bench_r02g_n$N.err; tail -1 gpurun_out/bench_r02g_n$N.log | cut -c1-400
