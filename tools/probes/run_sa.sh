run() { python tools/profile_step.py --skip-encoder --trace 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read())
print({'step': round(d['decode_step_ms'],3), 'by_pos': {k: round(v,2) for k,v in list(d['step_ms_by_position'].items())[::3]}, 'sa_end_us': round(d['eager_kernels_at_end']['decoder_self_attention']['avg_us'],1)})"; }
echo "default (8, 8)"; run
for f in tools/probes/libs/lib_sr*.so; do echo $f; B200W_LIB=$PWD/$f run; done
echo "default (8, 8)"; run
