"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA function: stall samples,
dynamic instructions, executed static footprint. Usage: python tools/ncu_by_function.py src.csv"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
agg = {}
linesrc = {}
infile = False; curfile = ''; cur = None
def toint(x):
  try: return int(x)
  except Exception: return 0
exec_static = {}
for r in rows:
  if not r: continue
  if r[0] in ('File Path', 'File Name'):
    infile = r[1].endswith('b200mj.cu'); curfile = r[1]; cur = None; continue
  if r[0] == 'Line No' or len(r) < 8: continue
  if r[0] != '':
    try: cur = int(r[0])
    except Exception: cur = None; continue
    if infile: linesrc[cur] = (r[1], toint(r[4]), toint(r[7]))
    else:
      a = agg.setdefault('LIB:' + curfile.split('/')[-1], [0, 0, 0]); a[0] += toint(r[4]); a[1] += toint(r[7])
  elif cur is not None and toint(r[7]) > 0:
    key = cur if infile else 'LIB:' + curfile.split('/')[-1]
    exec_static[key] = exec_static.get(key, 0) + 1
# full source listing (from `--print-source cuda`), optional 3rd arg; falls back to lines seen in the sass dump
full = {}
if len(sys.argv) > 3:
  ok = False
  for r in csv.reader(open(sys.argv[3])):
    if r and r[0] in ('File Path', 'File Name'): ok = r[1].endswith('b200mj.cu'); continue
    if ok and len(r) >= 2 and r[0].isdigit(): full[int(r[0])] = r[1]
else:
  full = {ln: v[0] for ln, v in linesrc.items()}
fn = '?'; fn_of = {}
for ln in sorted(full):
  txt = full[ln]
  m = re.match(r'^(?:extern "C" )?__(?:device|global)__[^(]*?(\w+)\(', txt)
  if m and not txt.strip().endswith(';'): fn = m.group(1)
  fn_of[ln] = fn
for ln, (txt, st, dy) in linesrc.items():
  a = agg.setdefault(fn_of.get(ln, '?'), [0, 0, 0]); a[0] += st; a[1] += dy
for k, v in exec_static.items():
  f = fn_of.get(k, k) if not isinstance(k, str) else k
  agg.setdefault(f, [0, 0, 0])[2] += v
ts = sum(a[0] for a in agg.values()); td = sum(a[1] for a in agg.values()); tx = sum(a[2] for a in agg.values())
print(f'total: stall samples {ts}, dynamic warp-instructions {td}, executed static {tx} ({tx*16/1024:.0f} KB)')
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
  print('%-30s stall %5.1f%%  dyn %5.1f%%  exec-static %5d (%5.1f KB)' % (k, 100 * a[0] / ts, 100 * a[1] / td, a[2], a[2] * 16 / 1024))
