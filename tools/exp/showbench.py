#!/usr/bin/env python
"""prints the headline fields of bench.py JSON lines: showbench.py a.json b.json ..."""
import sys, json
for f in sys.argv[1:]:
    try:
        for l in open(f):
            if l.startswith('{'):
                d = json.loads(l)
                print(f, d['config'].get('workload'), round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 3),
                      'verified', d.get('verified'), 'sm_mhz', d['clocks']['sm_mhz'], d['clocks']['reasons'])
    except Exception as e:
        print(f, 'ERR', e)
