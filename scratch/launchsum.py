import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; data = []
for r in rows:
    if r and r[0] == 'ID': hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(dict(zip(hdr, r)))
agg = collections.OrderedDict()
for d in data:
    name = d['Kernel Name'][:78]
    v = float(d['Metric Value'].replace(',', '')); unit = d['Metric Unit']
    if unit == 'ns': v /= 1000.0
    elif unit == 'ms': v *= 1000.0
    agg.setdefault((name, d['Grid Size'], d['Block Size']), []).append(v)
for (k, g, b), v in agg.items():
    print('%-80s grid=%-14s blk=%-12s n=%3d mean=%8.2f us min=%8.2f max=%8.2f' % (k, g, b, len(v), sum(v)/len(v), min(v), max(v)))
