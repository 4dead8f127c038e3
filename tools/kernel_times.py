import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hdr = next(r for r in rows if r[0] == 'ID')
ik, iv, ig, ib = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size'), hdr.index('Block Size')
agg = {}
for r in rows[rows.index(hdr) + 1:]:
  try:
    agg.setdefault((r[ik][:24], r[ig], r[ib]), []).append(float(r[iv].replace(',', '')))
  except Exception:
    pass
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
  print(k, len(v), 'mean ms %.3f  share %.1f%%' % (sum(v) / len(v) / 1e6, 100 * sum(v) / tot))
