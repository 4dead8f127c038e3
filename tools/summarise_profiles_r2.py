"""Turn gpurun_out/<tag>_{launches.csv, prof_acc.ncu-rep, prof_pos.ncu-rep} (tools/profile_bench_r2.sh) into the tracked
summaries under profiles/ (run here, no GPU).  Usage: python tools/summarise_profiles_r2.py r2c"""
import collections, csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
tag = sys.argv[1] if len(sys.argv) > 1 else 'r2c'
out = 'r2'
summary = {}

# ---- launch list: per kernel launches / time / DRAM bytes ---------------------------------------------------------
rows = [r for r in csv.reader(open(os.path.join(G, f'{tag}_launches.csv'))) if r]
hdr = next(r for r in rows if r and r[0] == 'ID')
ik, im, iv, iid = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
per = collections.defaultdict(lambda: collections.defaultdict(dict))
for r in rows[rows.index(hdr) + 1:]:
  try:
    name = re.sub(r'\(.*', '', r[ik]).replace('void ', '')
    per[name][r[iid]][r[im]] = float(r[iv].replace(',', ''))
  except Exception:
    pass
tot = sum(v.get('gpu__time_duration.sum', 0) for k in per.values() for v in k.values())
lines = ['kernel,launches,total_us,share_of_library_time,mean_us,dram_read_MB_per_launch,dram_write_MB_per_launch']
dram = {}
for k, v in sorted(per.items(), key=lambda kv: -sum(x.get('gpu__time_duration.sum', 0) for x in kv[1].values())):
  t = [x.get('gpu__time_duration.sum', 0) for x in v.values()]
  rd = [x.get('dram__bytes_read.sum', 0) for x in v.values()]; wr = [x.get('dram__bytes_write.sum', 0) for x in v.values()]
  lines.append(f'"{k}",{len(t)},{sum(t) / 1e3:.1f},{sum(t) / tot:.4f},{sum(t) / len(t) / 1e3:.1f},{sum(rd) / len(rd) / 1e6:.2f},{sum(wr) / len(wr) / 1e6:.2f}')
  dram[k] = (sum(rd) + sum(wr), len(t))
open(os.path.join(P, f'{out}_launches_summary.csv'), 'w').write('\n'.join(lines) + '\n')
# launches per control step of the humanoid workload: 2 groups x (4 pos + 5 x 4 acc buckets + 1 posfinal) = 50
nlaunch = sum(n for _, n in dram.values())
per_step = 2 * (4 + 5 * 4 + 1)
summary['dram_bytes_per_step'] = sum(b for b, _ in dram.values()) / nlaunch * per_step
summary['dram_bytes_per_step_by_kernel'] = {k: b / n * (per_step * n / nlaunch) for k, (b, n) in dram.items()}
summary['launches_in_list'] = nlaunch


def raw(rep):
  txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  r = list(csv.reader(txt.splitlines()))
  return dict(zip(r[0], r[2])), dict(zip(r[0], r[1]))


keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__shared_mem_per_block_dynamic', 'sm__icc_request_hit_rate.pct', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active']
for which in ('acc', 'pos'):
  rep = os.path.join(G, f'{tag}_prof_{which}.ncu-rep')
  if not os.path.exists(rep):
    continue
  d, unit = raw(rep)
  m = {'kernel': d.get('Kernel Name')}
  for k in keep:
    if k in d:
      m[k] = [d[k], unit.get(k, '')]
  stalls = {}
  for k, v in d.items():
    mm = re.match(r'smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio', k)
    if mm and not mm.group(1).endswith('not_issued'):
      try:
        stalls[mm.group(1)] = round(float(v), 3)
      except ValueError:
        pass
  m['stalls_per_issue'] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
  json.dump(m, open(os.path.join(P, f'{out}_{which}_kernel_metrics.json'), 'w'), indent=1)
  if which == 'acc':
    f = lambda k: float(d[k])
    summary['dominant_kernel'] = 'b200mj_acc_tn_kernel<false, 27>'
    summary['dominant_kernel_utilisation_pct'] = dict(
        issue_slots=round(f('smsp__issue_active.avg.pct_of_peak_sustained_active'), 2),
        lsu_pipe=round(f('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'), 2),
        fp64_pipe=round(f('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'), 2),
        alu_pipe=round(f('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'), 2),
        ipc_per_sm=round(f('sm__inst_executed.avg.per_cycle_active'), 2),
        warps_active=round(f('sm__warps_active.avg.pct_of_peak_sustained_active'), 2),
        icache_hit_rate=round(f('sm__icc_request_hit_rate.pct'), 2))
    summary['kernel_us_under_ncu'] = f('gpu__time_duration.sum')
  # per-function / opcode mix / SASS excerpt from the source page
  cs = os.path.join(G, f'{tag}_{which}_src_cs.csv'); c = os.path.join(G, f'{tag}_{which}_src_c.csv')
  open(cs, 'w').write(subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout)
  open(c, 'w').write(subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda'], capture_output=True, text=True).stdout)
  byfn = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_by_function.py'), cs, '28', c], capture_output=True, text=True).stdout
  open(os.path.join(P, f'{out}_{which}_by_function.txt'), 'w').write(byfn)
  sass = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
  srows = list(csv.reader(sass.splitlines()))
  h = next((r for r in srows if r and r[0] in ('Address', '#')), None)
  ops = collections.Counter(); total = 0; excerpt = []
  if h:
    isrc = h.index('Source'); iex = h.index('Instructions Executed') if 'Instructions Executed' in h else None
    for r in srows[srows.index(h) + 1:]:
      if len(r) <= isrc:
        continue
      ins = re.sub(r'^@!?U?P\d+\s+', '', r[isrc]).split()
      if not ins:
        continue
      n = int(r[iex]) if iex is not None and r[iex].isdigit() else 0
      ops[ins[0].split('.')[0]] += n; total += n
      excerpt.append((n, r[isrc]))
    mix = [f'dynamic SASS opcode mix of {d.get("Kernel Name", which)[:60]} ({total} warp-instructions in the captured launch)']
    for o, n in ops.most_common(22):
      mix.append(f'{o:12s} {100 * n / max(total, 1):5.1f} %')
    fp64 = sum(n for o, n in ops.items() if o in ('DFMA', 'DMUL', 'DADD', 'DSETP', 'MUFU'))
    mix.append(f'FP64 arithmetic (DFMA+DMUL+DADD+DSETP+MUFU): {100 * fp64 / max(total, 1):.1f} %')
    open(os.path.join(P, f'{out}_{which}_opcode_mix.txt'), 'w').write('\n'.join(mix) + '\n')
    if which == 'acc':
      # the hottest straight-line stretch: tn_factor's trailing update (LDS.128 broadcast + DFMA with register operands)
      best = max(range(len(excerpt) - 40), key=lambda i: sum(1 for n, t in excerpt[i:i + 40] if 'DFMA' in t and n > 0)) if len(excerpt) > 80 else 0
      open(os.path.join(P, f'{out}_acc_sass_excerpt.txt'), 'w').write(
          'SASS of the final compile-time-size acceleration kernel, 40 instructions inside tn_factor<27> (executed count, instruction):\n' +
          '\n'.join(f'{n:8d}  {t}' for n, t in excerpt[best:best + 40]) + '\n')
summary['source'] = f'profiles/{out}_*: ncu captures of `python bench.py --steps 2..3 --warmup 3 --no-cpu --no-configs` (tools/profile_bench_r2.sh {tag})'
json.dump(summary, open(os.path.join(P, 'summary.json'), 'w'), indent=1)
print(json.dumps(summary, indent=1)[:1500])
