python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02e.log 2>&1; tail -3 gpurun_out/pytest_r02e.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02e.log 2> gpurun_out/bench_r02e.err; tail -c 600 gpurun_out/bench_r02e.err
python tools/ncu_cross.py > gpurun_out/ncu_cross_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cross_attention -s 1 -c 2 -f -o gpurun_out/cross_ring_r02 python tools/ncu_cross.py > gpurun_out/ncu_cross_ring.log 2>&1; tail -2 gpurun_out/ncu_cross_ring.log
