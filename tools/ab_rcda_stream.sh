for v in 0 1 2 3; do
CDETR_RCDA_STREAM=$v timeout 300 python bench.py --skip-cpu --skip-matcher --no-full 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('stream=$v', round(d['ms_per_step'],3), d['loss'], {k:v['ms'] for k,v in d['roofline_attention']['per_kernel'].items() if 'rcda' in k})"
done
