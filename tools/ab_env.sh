#!/bin/bash
# A/B of an environment switch on one box: tools/ab_env.sh VAR "v1 v2 ..." [repeats]  -> ms/step of the C3 bench per value
VAR=$1; VALS=$2; REP=${3:-2}
for r in $(seq $REP); do for v in $VALS; do
env $VAR=$v timeout 300 python bench.py --skip-cpu --skip-matcher --no-full 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$VAR=$v', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['loss'])"
done; done
