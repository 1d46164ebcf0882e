#!/bin/bash
# tools/knob_sweep.sh "ENV=VAL ENV2=VAL" ...  — one bench run per argument, prints value + per-class ms (timing experiments)
for cfg in "$@"; do
  env $cfg timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /tmp/ks.json 2>/tmp/ks.err
  python - "$cfg" <<'PY'
import json,sys
try:
    d=json.load(open('/tmp/ks.json'))
    c=d['conv_classes']
    print(sys.argv[1], '| %.0fM'%(d['value']/1e6), ' '.join('%s=%.2f'%(k,v['ms']) for k,v in c.items()))
except Exception as e:
    print(sys.argv[1], 'FAILED', e, open('/tmp/ks.err').read()[-300:])
PY
done
