#!/bin/bash
# Which clock sampler disturbs the timed loop?  per-step times of bench.py's resident region with the NVML thread,
# the nvidia-smi process and no sampler.   tools/bench_outliers.sh  (on a GPU box)
for mode in nvml smi none; do
  for rep in 1 2; do
    if [ $mode = none ]; then export NAVC_NO_SAMPLER=1; else unset NAVC_NO_SAMPLER; export NAVC_SAMPLER=$mode; fi
    timeout 400 python bench.py --steps 40 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r = sorted(d['per_step_ms']['resident']); e = sorted(d['per_step_ms']['e2e'])
print('$mode', 'value %.0f e2e %.0f' % (d['value'], d['e2e']['value']), 'resident median %.2f max %.2f n>20ms %d' % (r[len(r)//2], r[-1], sum(x > 20 for x in r)),
      'e2e median %.2f max %.2f n>20ms %d' % (e[len(e)//2], e[-1], sum(x > 20 for x in e)), d['clocks'].get('samples'), d['clocks'].get('sampler'))
"
  done
done
