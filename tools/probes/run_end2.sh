python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02k.log 2>&1; tail -3 gpurun_out/pytest_r02k.log
python tools/ncu_small.py 5 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_small_mma -s 2 -c 1 -f -o gpurun_out/small_mma_r02 python tools/ncu_small.py 5 > gpurun_out/ncu_small_mma.log 2>&1; tail -1 gpurun_out/ncu_small_mma.log
