python tools/time_cross.py
for f in tools/probes/libs/lib_r*.so; do B200W_LIB=$PWD/$f python tools/time_cross.py; done
B200W_CROSS_STREAM=0 python tools/time_cross.py
for f in tools/probes/libs/lib_r8_c5.so tools/probes/libs/lib_r6_c6.so; do echo $f; B200W_LIB=$PWD/$f python tools/profile_step.py --skip-encoder; done
