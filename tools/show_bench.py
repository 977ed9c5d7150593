#!/usr/bin/env python
"""Print the interesting parts of a bench.py JSON line. Usage: show_bench.py file.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
print('value', round(d['value'], 1), 'mb/s  ms/step', round(d['ms_per_step'], 3), ' e2e', round(d['e2e']['value'], 1),
      ' gather_gbs', round(d['gather_gbs'], 1), ' launches', d['gpu_launches'], ' setup_s', d.get('setup_s'))
print(' roofline', d['roofline']['kernel'], round(d['roofline']['achieved'], 1), round(d['roofline']['frac'], 3), 'clocks', d['clocks'])
print(' cpu', d['cpu_baseline'] and (round(d['cpu_baseline']['value'], 2), d['cpu_baseline']['stage_s']))
for m, r in d['variants'].items():
    for k in ('value', 'e2e'):
        x = r[k]
        print(' ', m, k, 'mb/s', round(x['minibatches_per_s'], 1), 'ms', round(x['ms_per_step'], 3), 'wall', round(x['wall_ms_per_step'], 3),
              'hit', round(x['hit_rate'], 3), 'launches', x['launches'])
        for kn, kv in x['kernels'].items():
            print('       %-58s avg_ms %.4f  (in pipeline %s)  GB/s %8.1f  frac %.3f (%s)' % (
                kn, kv['avg_ms'], ('%.4f' % kv['avg_ms_in_pipeline']) if 'avg_ms_in_pipeline' in kv else '-', kv['achieved_gbs'], kv['frac'], kv['bound']))
        if x.get('gather'):
            g = x['gather']
            print('     gather: %.1f GB/s payload, %.3f ms/batch, hit %.3f; hit kernel frac %.3f' % (g['gather_gbs'], g['ms_per_batch'], g['hit_rate'], g['hit']['frac']),
                  ('miss %.1f GB/s frac %.3f' % (g['miss']['achieved_gbs'], g['miss']['frac'])) if 'miss' in g else '')
