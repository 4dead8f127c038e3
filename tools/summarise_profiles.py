"""Turn gpurun_out/{launches.csv, prof_bench.ncu-rep} into the tracked summaries under profiles/ (run here, no GPU)."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'
os.makedirs(P, exist_ok=True)
summary = {}
# ---- launch list -------------------------------------------------------------------------------------------
lc = os.path.join(G, 'launches.csv')
if os.path.exists(lc):
  rows = [r for r in csv.reader(open(lc)) if r]
  hdr = next(r for r in rows if r and r[0] == 'ID')
  ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
  per = {}
  for r in rows[rows.index(hdr) + 1:]:
    try: per.setdefault(r[ik], []).append(float(r[iv].replace(',', '')))
    except Exception: pass
  tot = sum(sum(v) for v in per.values())
  lines = ['kernel,launches,total_ns,share']
  for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
    lines.append(f'"{k[:90]}",{len(v)},{sum(v):.0f},{sum(v) / tot:.4f}')
  open(os.path.join(P, f'{tag}_launches_summary.csv'), 'w').write('\n'.join(lines) + '\n')
  step = [k for k in per if 'b200mj_step_kernel' in k]
  if step:
    summary['step_kernel_share_of_gpu_time_under_ncu'] = sum(per[step[0]]) / tot
    summary['step_kernel_launches_in_list'] = len(per[step[0]])
# ---- full capture ---------------------------------------------------------------------------------------------
rep = os.path.join(G, 'prof_bench.ncu-rep')
if os.path.exists(rep):
  raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(raw.splitlines()))
  d = dict(zip(rows[0], rows[2]))
  keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
          'sm__inst_executed.avg.per_cycle_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
          'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem',
          'sm__icc_request_hit_rate.pct', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
          'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
          'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active']
  unit = dict(zip(rows[0], rows[1]))
  metrics = {k: (d.get(k), unit.get(k)) for k in keep if k in d}
  stalls = {k.split('issue_stalled_')[1]: float(d[k]) for k in rows[0] if 'pcsamp_warps_issue_stalled' in k and 'not_issued' not in k}
  ts = sum(stalls.values()) or 1
  metrics['stall_mix_pct'] = {k: round(100 * v / ts, 2) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]}
  json.dump(metrics, open(os.path.join(P, f'{tag}_step_kernel_metrics.json'), 'w'), indent=1)
  def num(k):
    v, u = d.get(k, '0'), unit.get(k, '')
    f = float(v.replace(',', ''))
    return f * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}.get(u, 1)
  summary['dram_bytes_per_launch'] = num('dram__bytes_read.sum') + num('dram__bytes_write.sum')
  # what the dominant kernel is actually bound by (it is nowhere near the HBM roofline): issue slots and pipes
  def pct(k):
    try: return round(float(d[k].replace(',', '')), 2)
    except Exception: return None
  summary['dominant_kernel_utilisation_pct'] = dict(
      issue_slots=pct('smsp__issue_active.avg.pct_of_peak_sustained_active'), lsu_pipe=pct('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'),
      fp64_pipe=pct('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'), alu_pipe=pct('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'),
      fma_pipe=pct('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'), ipc_per_sm=pct('sm__inst_executed.avg.per_cycle_active'),
      warps_active=pct('sm__warps_active.avg.pct_of_peak_sustained_active'))
  summary['kernel_ms_under_ncu'] = float(d['gpu__time_duration.sum'].replace(',', '')) * {'ms': 1, 'us': 1e-3, 'ns': 1e-6, 'msecond': 1, 'usecond': 1e-3, 'nsecond': 1e-6}.get(unit['gpu__time_duration.sum'], 1)
  src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
  open(os.path.join(G, 'bench_src_cs.csv'), 'w').write(src)
  srcc = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda', '--csv'], capture_output=True, text=True).stdout
  open(os.path.join(G, 'bench_src_c.csv'), 'w').write(srcc)
  by = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_by_function.py'), os.path.join(G, 'bench_src_cs.csv'), '40',
                       os.path.join(G, 'bench_src_c.csv')], capture_output=True, text=True).stdout
  open(os.path.join(P, f'{tag}_step_kernel_by_function.txt'), 'w').write(by)
# ---- DRAM traffic of the whole step group --------------------------------------------------------------------
dc = os.path.join(G, 'step_dram.csv')
if os.path.exists(dc):
  rows = [r for r in csv.reader(open(dc)) if r]
  hdr = next(r for r in rows if r and r[0] == 'ID')
  ik, im, iu, iv = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Unit'), hdr.index('Metric Value')
  tot = 0.0; fused = 0; per_kernel = {}
  for r in rows[rows.index(hdr) + 1:]:
    if 'dram__bytes' in r[im]:
      v = float(r[iv].replace(',', '')) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}.get(r[iu], 1)
      tot += v; per_kernel[r[ik][:24]] = per_kernel.get(r[ik][:24], 0) + v
    if r[im] == 'gpu__time_duration.sum' and 'b200mj_posfinal_kernel' in r[ik]: fused += 1
  if fused:
    summary['dram_bytes_per_step'] = tot / fused
    summary['dram_bytes_per_step_by_kernel'] = {k: v / fused for k, v in per_kernel.items()}
    summary['dram_steps_in_capture'] = fused
lc2 = os.path.join(P, f'{tag}_launches_summary.csv')
if os.path.exists(lc2):
  rows = list(csv.reader(open(lc2)))[1:]
  mine = [r for r in rows if r and 'b200mj' in r[0]]
  if mine: summary['dominant_kernel'] = max(mine, key=lambda r: float(r[2]))[0]
summary['source'] = f'profiles/{tag}_*: ncu captures of `python bench.py --steps 2..3 --warmup 3 --no-cpu` (tools/profile_bench.sh)'
json.dump(summary, open(os.path.join(P, 'summary.json'), 'w'), indent=1)
print(json.dumps(summary, indent=1))
