for b in 118 120 122 133 148 150; do python tools/time_cross.py $b; done
