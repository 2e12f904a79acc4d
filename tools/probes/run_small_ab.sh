for r in 1 2; do for v in a b c; do B200W_LIB=$PWD/tools/probes/libs/lib_$v.so python tools/probes/run_small_ab.py 2>&1 | tail -1; done; done
