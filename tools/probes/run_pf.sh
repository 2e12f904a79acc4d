for n in 0 4 8 12 16; do
  echo "== B200W_CROSS_PF=$n"
  B200W_CROSS_PF=$n python tools/profile_step.py --skip-encoder --eager 2 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().split('\n')[-1] if False else sys.stdin.read())
k = d.get('eager_kernels_ms_per_step', {})
print(json.dumps({'decode_step_ms': d.get('decode_step_ms'), 'cross_us': k.get('decoder_cross_attention', {}).get('avg_us'), 'chain_us': k.get('dec_chain', {}).get('avg_us'), 'sa_us': k.get('decoder_self_attention', {}).get('avg_us')}))"
done
