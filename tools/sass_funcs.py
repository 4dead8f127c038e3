"""Split one kernel of a cuobjdump -sass dump into its subroutines (at RET) and print size + opcode mix of each.
Usage: python tools/sass_funcs.py build/new.sass <kernel-name-prefix> [min_instr]"""
import re, collections, sys
txt = open(sys.argv[1]).read()
funcs = re.split(r'\n\s*Function : ', txt)
f = [x for x in funcs[1:] if x.startswith(sys.argv[2])][0]
minn = int(sys.argv[3]) if len(sys.argv) > 3 else 100
ops = []
for l in f.split('\n'):
  m = re.search(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
  if m: ops.append((int(m.group(1), 16), m.group(2)))
strip = lambda o: re.sub(r'^@!?U?P\d+\s+', '', o)
rets = [a for a, o in ops if strip(o).startswith('RET')]
print(len(ops), 'instructions;', len(rets), 'RETs')
prev = -1
for r in rets + [ops[-1][0]]:
  seg = [o for a, o in ops if prev < a <= r]
  c = collections.Counter(strip(o).split()[0].split('.')[0] for o in seg)
  if len(seg) >= minn: print(hex(prev + 16), hex(r), len(seg), c.most_common(14))
  prev = r
