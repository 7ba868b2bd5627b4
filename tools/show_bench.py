#!/usr/bin/env python
"""Print the figures of one or more bench.py JSON lines (files) that matter when comparing runs."""
import json, sys
for f in sys.argv[1:]:
    try:
        txt = open(f).read()
        d = json.loads([ln for ln in txt.splitlines() if ln.startswith("{")][-1])
    except Exception as e:
        print(f, 'ERR', e, txt[-800:]); continue
    print('==', f)
    for k in ('value', 'ms_per_step', 'replays_ms_per_step', 'gpu_launches', 'clocks', 'exchange_overhead_us', 'exchange_wait_per_rank', 'exchange_parity'):
        if d.get(k) is not None: print(' ', k, d.get(k))
    r = d.get('roofline')
    if r: print('  roofline', {k: r.get(k) for k in ('achieved', 'frac', 'kernel', 'plain_kernel_ms_event_pairs_mean', 'plain_kernel_ms_min')})
    e = d.get('e2e')
    if e: print('  e2e', {k: e.get(k) for k in ('value', 'ms_per_step', 'h2d_bytes_per_step')})
    if d.get('configs4'): print('  c4', json.dumps(d['configs4'])[:1600])
    if d.get('extra'):
        for k, v in d['extra'].get('configs', d['extra']).items(): print('  extra', k, json.dumps(v)[:300])
