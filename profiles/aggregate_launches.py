"""Aggregate an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list
per kernel: launches, summed duration, share of the kernel time, DRAM bytes.  Usage: aggregate_launches.py X.csv"""
import collections
import csv
import re
import sys

lines = open(sys.argv[1]).readlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
agg = collections.defaultdict(lambda: collections.defaultdict(float))
ids = collections.defaultdict(set)
for row in csv.DictReader(lines[start:]):
    k = re.sub(r'\(.*', '', row['Kernel Name'])
    v = float(row['Metric Value'].replace(',', ''))
    u, m = row['Metric Unit'], row['Metric Name']
    if m == 'gpu__time_duration.sum':
        v *= {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3}[u]
    else:
        v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
    agg[k][m] += v
    ids[k].add(row['ID'])
tot = sum(a['gpu__time_duration.sum'] for a in agg.values())
conv = 0.0
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['gpu__time_duration.sum']):
    dram = a['dram__bytes_read.sum'] + a['dram__bytes_write.sum']
    if 'conv_tc' in k:
        conv += dram
    print('%-40s n=%4d t=%9.1f us share=%.3f dram=%.3f GB' % (k[:40], len(ids[k]), a['gpu__time_duration.sum'],
                                                             a['gpu__time_duration.sum'] / tot, dram / 1e9))
print('total %.1f us; tensor-core conv DRAM traffic %.3f GB' % (tot, conv / 1e9))
