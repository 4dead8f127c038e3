"""bench.py — env-steps/s of suite.humanoid:run, batch 8192 per GPU, random-action rollout (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the CPU restatement of the reference path, host cores

One "step" = one `Environment.step` for the whole batch = n_sub_steps(5) physics steps + reward + observation
(reference: rl/control.py:99-127, suite/humanoid.py:30). Prints ONE JSON line on rank 0. See DESIGN.md §Measurement.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'env-steps/sec suite.humanoid:run batch 8192 (per GPU) random-action rollout'
UNIT = 'env-steps/s'
BATCH = 8192
NSUB = 5
# compulsory fp64 bytes per humanoid env-step (5 fused substeps + observation-contract outputs)
ALGO_BYTES_PER_ENV_STEP = 4068
OBS_DIM = 67


def _peaks():
  p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
  return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
  """SM clock + clocks-event (throttle) reasons sampled DURING the timed region.

  Samples are taken inline from the benchmark thread through NVML every few steps (0.15 ms per sample). A polling
  child process (`nvidia-smi -lms 200`) or a polling thread was measured to cost this workload 15-40 %: the step
  issues ~17 launches plus cross-stream events, and concurrent driver queries stall them. nvidia-smi is the fallback
  (one query per sample) when pynvml is unavailable.
  """
  BITS = dict(sw_power_cap=0x4, hw_slowdown=0x8, sw_thermal_slowdown=0x20, hw_thermal_slowdown=0x40)

  def __init__(self, index, uuid=None):
    self.index, self.rows, self.mode = index, [], None
    try:
      import pynvml
      pynvml.nvmlInit()
      h = None
      if uuid:
        for cand in (f'GPU-{uuid}', str(uuid)):
          try:
            h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode())
            break
          except Exception:
            h = None
      self._h = h if h is not None else pynvml.nvmlDeviceGetHandleByIndex(index)
      self._nv = pynvml
      self._max = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
      self._reasons = getattr(pynvml, 'nvmlDeviceGetCurrentClocksEventReasons', None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
      self.mode = 'nvml-inline'
    except Exception:
      self.mode = 'nvidia-smi-inline'

  def sample(self):
    try:
      if self.mode == 'nvml-inline':
        sm = self._nv.nvmlDeviceGetClockInfo(self._h, self._nv.NVML_CLOCK_SM)
        bits = int(self._reasons(self._h))
        self.rows.append((float(sm), float(self._max), [n for n, b in self.BITS.items() if bits & b]))
      else:
        out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=clocks.sm,clocks.max.sm,'
                              'clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,'
                              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.hw_thermal_slowdown',
                              '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=10).stdout
        f = [x.strip() for x in out.strip().split(',')]
        self.rows.append((float(f[0]), float(f[1]), [n for n, v in zip(self.BITS, f[2:6]) if v.lower().startswith('active')]))
    except Exception:
      pass

  def summary(self):
    sm = [r[0] for r in self.rows]
    return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max((r[1] for r in self.rows), default=None),
                reasons=sorted({n for r in self.rows for n in r[2]}), samples=len(sm), source=self.mode)


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path on the host cores
# ------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
  """One host process: builds its environments, waits at the shared start time, rolls them out in C."""
  t, nenv, warmup_steps, timed_steps, start_at = args
  import numpy as np
  from dm_control_b200 import testing_models as tm
  from oracle import oracle as om
  model = tm.load('humanoid')
  q0, v0 = tm.initial_states(model, 'humanoid', nenv, 7000 + t)
  envs = []
  for e in range(nenv):
    o = om.OraclePhysics(model)
    o.qpos[:] = q0[e]; o.qvel[:] = v0[e]; o.forward()
    envs.append(o)
  tape = np.random.RandomState(100 + t).uniform(-1, 1, (warmup_steps + timed_steps, nenv, model.nu))
  for j, o in enumerate(envs):
    o.rollout(tape[:warmup_steps, j], NSUB)
  while time.time() < start_at:
    time.sleep(0.001)
  t0 = time.time()
  for j, o in enumerate(envs):
    o.rollout(tape[warmup_steps:, j], NSUB)
  return t0, time.time()


def time_cpu(warmup_steps=5, timed_steps=150, nenv_per_proc=8, procs=None):
  """Times `control_step(5)` (legacy ordering, engine.py:147-162) of the scalar CPU oracle on all host cores.

  One forked process per core, each owning `nenv_per_proc` environments and stepping them inside one C call;
  all processes start their timed rollouts at the same wall-clock instant and the slowest sets the time.
  Sized to ~10-30 s of CPU work in total. Returns (env_steps_per_s, cores, sample description, ms per env-step).
  """
  import multiprocessing as mp
  from oracle import oracle as om
  om.build()
  procs = procs or os.cpu_count() or 1
  ctx = mp.get_context('fork')
  start_at = time.time() + 3.0 + 0.02 * procs          # leave time for every worker to build + warm up
  with ctx.Pool(procs) as pool:
    spans = pool.map(_cpu_worker, [(t, nenv_per_proc, warmup_steps, timed_steps, start_at) for t in range(procs)], chunksize=1)
  dt = max(b for _, b in spans) - min(a for a, _ in spans)
  nenv = procs * nenv_per_proc
  value = nenv * timed_steps / dt
  sample = (f'{nenv} envs ({nenv_per_proc}/process x {procs} processes) x {timed_steps} env-steps after {warmup_steps} '
            f'warm-up, seeded humanoid:run states, uniform(-1,1) actions')
  return value, procs, sample, dt * 1e3 / timed_steps


def cpu_sample_sizes(steps, warmup):
  """The CPU arm's bounded sample of the GPU arm's workload: the same untimed settle (warmup + 1 env-steps from the
  same family of seeded states, so the timed window sees the same contact load), the same number of timed env-steps
  (clamped to 20..300), and enough environments per process for a few seconds of work on every core."""
  warm = min(max(warmup, 3) + 1, 60)
  timed = max(20, min(steps, 300))
  nenv = max(4, min(48, 1600 // timed))
  return warm, timed, nenv


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  warm, timed, nenv = cpu_sample_sizes(max(1, args.steps), args.warmup)
  kind, note = 'port', 'restated CPU oracle (oracle/mjoracle.cpp), NOT libmujoco: MuJoCo is absent from this image'
  value = None
  try:
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import time_mujoco_cpu
    if time_mujoco_cpu.available():      # a machine that has the real reference: time that instead
      value, cores, sample, ms = time_mujoco_cpu.time_reference(warm, timed, max(1, nenv // 4))
      kind, note = 'reference', 'unmodified dm_control suite.load(humanoid, run) on mujoco, Environment.step'
  except Exception as ex:
    sys.stderr.write(f'real-reference arm failed ({ex!r}); timing the oracle port\n')
    value = None
  if value is None:
    value, cores, sample, ms = time_cpu(warm, timed, nenv)
  line = dict(impl='reference', metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
              ms_per_step=ms, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64', data='synthetic',
              config=dict(workload='suite.humanoid:run, 5 physics substeps per env-step, random actions', note=note),
              cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind=kind, sample=sample),
              e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
  print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
def run_gpu(args):
  import torch
  import torch.distributed as dist
  from dm_control_b200 import lib as blib
  from dm_control_b200 import suite
  from dm_control_b200 import testing_models as tm

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
  L = blib.load()

  env = suite.load('humanoid', 'run', batch=BATCH, seed=1000 + rank, device=dev)
  phys = env.physics
  phys.check_errors = False          # no device->host sync inside the rollout; warnings are summed at the end
  model = phys.model
  # seeded, partly-in-contact start states (same family the parity tests use), then the task's own settle
  q0, v0 = tm.initial_states(model, 'humanoid', BATCH, seed=rank)
  phys.data.qpos.copy_(torch.as_tensor(q0, device=dev)); phys.data.qvel.copy_(torch.as_tensor(v0, device=dev))
  phys.forward()
  env._reset_next.zero_()
  # the task's ~80 tiny reward/observation launches replay as one CUDA graph (falls back to eager if capture fails)
  env._graph_task_ops = not os.environ.get('B200_BENCH_NO_GRAPH')
  gen = torch.Generator(device=dev).manual_seed(1234 + rank)
  actions = torch.empty(BATCH, model.nu, dtype=torch.float64, device=dev)
  packed = torch.empty(BATCH, OBS_DIM + 2, dtype=torch.float64, device=dev)
  from dm_control_b200 import sharding
  gathered = torch.empty(BATCH * world, OBS_DIM + 2, dtype=torch.float64, device=dev) if (world > 1 and rank == 0) else None
  flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)   # > 126 MB L2

  def pack(ts):
    o = ts.observation
    packed[:, :21] = o['joint_angles']; packed[:, 21] = o['head_height']; packed[:, 22:34] = o['extremities']
    packed[:, 34:37] = o['torso_vertical']; packed[:, 37:40] = o['com_velocity']; packed[:, 40:67] = o['velocity']
    packed[:, 67] = ts.reward; packed[:, 68] = ts.discount
    if world > 1:
      # NCCL over NVLink: observations/rewards to rank 0, in environment order (north_star)
      sharding.gather_to_rank0(packed, BATCH * world, out=gathered)

  def one_step():
    flush.fill_(0.0)                                   # L2 flush between timed iterations (inside the timed region)
    actions.uniform_(-1, 1, generator=gen)
    pack(env.step(actions))

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  sampler = ClockSampler(local, getattr(torch.cuda.get_device_properties(local), "uuid", None)) if (rank == 0 and not os.environ.get("B200_BENCH_NO_SAMPLER")) else None
  # settle to the steady-state contact load the metric is quoted on (20 warm-up env-steps minimum)
  try:
    one_step()
  except Exception as ex:           # graph capture unsupported for some op: eager task ops
    if not env._graph_task_ops:
      raise
    sys.stderr.write(f'task-op graph capture failed ({ex!r}); running the task ops eagerly\n')
    env._graph_task_ops = False; env._graph = None
    torch.cuda.synchronize()
    one_step()
  for _ in range(max(args.warmup, 3)):
    one_step()
  barrier()

  # ---- device-resident arm -------------------------------------------------------------------------------
  launches0 = L.b200mj_launch_count()
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
  kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
  barrier()
  ev[0].record()
  every = max(1, args.steps // 8)
  for i in range(args.steps):
    if sampler and i % every == every // 2:
      sampler.sample()                                 # clocks / throttle reasons while the region is running
    flush.fill_(0.0)
    actions.uniform_(-1, 1, generator=gen)
    env._task.before_step(actions, phys)
    kev[i][0].record()
    phys.step(env.n_sub_steps)                         # one b200mj_step call: the step's whole kernel group
    kev[i][1].record()
    env._task.after_step(phys)
    reward, obs = env._reward_and_observation()
    pack(type('TS', (), dict(observation=obs, reward=reward, discount=torch.ones_like(reward))))
  ev[1].record()
  barrier()
  ms_total = ev[0].elapsed_time(ev[1])
  kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
  launches = L.b200mj_launch_count() - launches0
  clocks = sampler.summary() if sampler else None
  t = torch.tensor([ms_total, kernel_ms], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  ms_total, kernel_ms = float(t[0]), float(t[1])
  value = BATCH * world * args.steps / (ms_total * 1e-3)

  # ---- end-to-end arm: HOST action buffer in, HOST observation/reward buffer out, every step -----------------
  # host action tape, drawn before the clock starts (drawing 172k doubles on one host core costs ~0.4 ms per step and
  # is the synthetic policy's time, not the path's); 16 pinned blocks, cycled
  cpu_gen = torch.Generator().manual_seed(77 + rank)
  act_tape = [torch.empty(BATCH, model.nu, dtype=torch.float64).uniform_(-1, 1, generator=cpu_gen).pin_memory() for _ in range(16)]
  out_host = torch.empty((BATCH * world if rank == 0 else BATCH), OBS_DIM + 2, dtype=torch.float64).pin_memory()
  barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for i in range(args.steps):
    flush.fill_(0.0)
    actions.copy_(act_tape[i % len(act_tape)], non_blocking=True)     # H2D from pinned host memory
    pack(env.step(actions))
    if rank == 0 and world > 1:
      out_host.copy_(gathered, non_blocking=True)                     # D2H of the gathered block
    else:
      out_host[:BATCH].copy_(packed, non_blocking=True)               # D2H
    torch.cuda.current_stream().synchronize()                         # the user reads obs before the next action
  e1.record()
  barrier()
  e2e_ms = e0.elapsed_time(e1)
  t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  e2e_value = BATCH * world * args.steps / (float(t[0]) * 1e-3)
  warn = phys.data.warning.sum(0)
  if world > 1:
    dist.all_reduce(warn)

  if rank == 0:
    peak, peak_src = _peaks()
    achieved = ALGO_BYTES_PER_ENV_STEP * BATCH / (kernel_ms * 1e-3) / 1e9
    prof = {}
    pj = os.path.join(ROOT, 'profiles', 'summary.json')
    if os.path.exists(pj):
      prof = json.load(open(pj))
    cpu = None
    if world == 1 and not args.no_cpu:
      v, cores, sample, _ = time_cpu(*cpu_sample_sizes(args.steps, args.warmup))
      cpu = dict(value=v, unit=UNIT, cores=cores, kind='port', sample=sample)
    line = dict(
        metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=ms_total / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64',
        data='synthetic',
        config=dict(workload='suite.humanoid:run', batch_per_gpu=BATCH, global_batch=BATCH * world, n_sub_steps=NSUB,
                    physics_steps_per_s=value * NSUB, parallelism=f'env-sharded x{world}',
                    actions='uniform(-1,1) generated on device', l2='256 MB flush write between steps, inside the timed region',
                    task_ops='one CUDA-graph replay' if env._graph_task_ops else 'eager torch ops',
                    obs_gather='NCCL gather of [B,69] f64 to rank 0 each step' if world > 1 else 'n/a (1 GPU)',
                    kernels=phys.describe(),
                    nconmax=model.nconmax, njmax=model.njmax),
        e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=BATCH * model.nu * 8,
                 d2h_bytes_per_step=BATCH * (OBS_DIM + 2) * 8 * (world if world > 1 else 1)),
        gpu_launches=int(launches),
        roofline=dict(bound='hbm', achieved=achieved, peak=peak, unit='GB/s', frac=achieved / peak,
                      traffic=prof.get('dram_bytes_per_step', prof.get('dram_bytes_per_launch')), peak_source=peak_src,
                      kernel='b200mj step group: [pos_kernel, acc_kernel x row-buckets] x (n_sub_steps-1), [pos_kernel, acclast_kernel x row-buckets], posfinal_kernel',
                      dominant_kernel=prof.get('dominant_kernel', 'b200mj_acc_kernel'),
                      dominant_kernel_utilisation_pct=prof.get('dominant_kernel_utilisation_pct'),
                      kernel_ms=kernel_ms, kernel_share_of_step=kernel_ms / (ms_total / args.steps),
                      algorithmic_bytes_per_launch=ALGO_BYTES_PER_ENV_STEP * BATCH,
                      note='latency/issue-bound fp64 kernel: compulsory traffic is ~4 kB per env-step, see DESIGN.md'),
        clocks=clocks, warnings=[int(x) for x in warn.tolist()])
    if cpu is not None:
      line['cpu_baseline'] = cpu
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=50)
  ap.add_argument('--warmup', type=int, default=20)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
  args = ap.parse_args()
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_gpu(args)


if __name__ == '__main__':
  main()
