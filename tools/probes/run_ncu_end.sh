python tools/profile_step.py --skip-encoder --eager 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"decode_chain|filter_argmax" -s 70 -c 4 -f -o gpurun_out/chain_fa_r02_end python tools/profile_step.py --skip-encoder --eager 1 > gpurun_out/ncu_chain_end.log 2>&1; tail -1 gpurun_out/ncu_chain_end.log
python tools/ncu_small.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_small -s 2 -c 1 -f -o gpurun_out/small_r02_end python tools/ncu_small.py > gpurun_out/ncu_small_end.log 2>&1; tail -1 gpurun_out/ncu_small_end.log
