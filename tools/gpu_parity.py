"""GPU-vs-oracle parity probe (diagnostic twin of tests/test_gpu_parity.py). Run under gpurun."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dm_control_b200 import mjcf_compile as mc
from dm_control_b200.physics import BatchedPhysics
from dm_control_b200 import testing_models as tm
from oracle.oracle import OraclePhysics

def relerr(a, b):
  a = np.asarray(a); b = np.asarray(b)
  return float(np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b)))) if a.size else 0.0

def run(name, B, ncontrol, nsub, seed=0):
  model = tm.load(name)
  q0, v0 = tm.initial_states(model, name, B, seed)
  rs = np.random.RandomState(seed + 1)
  tape = rs.uniform(-1, 1, (ncontrol, B, model.nu))
  phys = BatchedPhysics(model, batch=B)
  phys.data.qpos.copy_(torch.as_tensor(q0)); phys.data.qvel.copy_(torch.as_tensor(v0))
  phys.forward()
  oracles = []
  for e in range(B):
    o = OraclePhysics(model); o.qpos[:] = q0[e]; o.qvel[:] = v0[e]; o.forward(); oracles.append(o)
  # forward-stage field comparison
  rep = {}
  for f in ('xpos', 'xmat', 'subtree_com', 'qfrc_bias', 'qfrc_passive', 'qacc', 'qfrc_constraint', 'sensordata', 'cvel', 'geom_xpos', 'site_xpos'):
    g = getattr(phys.data, f).cpu().numpy().reshape(B, -1)
    o = np.stack([np.asarray(getattr(oo, f)).reshape(-1) for oo in oracles])
    rep[f] = relerr(g, o)
  gM = phys.data.qM.cpu().numpy(); oM = np.stack([oo.M_dense() for oo in oracles]); rep['qM'] = relerr(gM, oM)
  gn = phys.data.ncon.cpu().numpy(); on = np.array([oo.ncon for oo in oracles]); rep['ncon_mismatch'] = int((gn != on).sum())
  ge = phys.data.nefc.cpu().numpy(); oe = np.array([oo.nefc for oo in oracles]); rep['nefc_mismatch'] = int((ge != oe).sum())
  print(name, 'forward:', {k: (float('%.2e' % v) if isinstance(v, float) else v) for k, v in rep.items()}, 'ncon max', int(on.max()), 'nefc max', int(oe.max()), flush=True)
  worst_q = worst_v = 0.0; pair_bad = 0; ncon_bad = 0; first_bad = None
  for t in range(ncontrol):
    phys.set_control(torch.as_tensor(tape[t]))
    phys.step(nsub)
    gq = phys.data.qpos.cpu().numpy(); gv = phys.data.qvel.cpu().numpy()
    gn = phys.data.ncon.cpu().numpy(); gg = phys.data.contact_geom.cpu().numpy()
    for e, o in enumerate(oracles):
      o.ctrl[:] = tape[t, e]; o.control_step(nsub)
    oq = np.stack([o.qpos for o in oracles]); ov = np.stack([o.qvel for o in oracles])
    eq, ev = relerr(gq, oq), relerr(gv, ov)
    worst_q, worst_v = max(worst_q, eq), max(worst_v, ev)
    for e, o in enumerate(oracles):
      cs = o.contact
      if len(cs) != gn[e]: ncon_bad += 1
      elif any((c.geom1, c.geom2) != tuple(gg[e, i]) for i, c in enumerate(cs)): pair_bad += 1
    if first_bad is None and max(eq, ev) > 1e-5: first_bad = t
  warn = phys.data.warning.cpu().numpy().sum(0)
  print(f'{name}: B={B} steps={ncontrol}x{nsub} rel-err qpos {worst_q:.2e} qvel {worst_v:.2e}  ncon-mismatch {ncon_bad} pair-mismatch {pair_bad} first>1e-5 at {first_bad} warnings {warn.tolist()} ws_bytes {phys.workspace_bytes()} epb {phys.envs_per_block()}', flush=True)
  return max(worst_q, worst_v)

def bench(name, B, nsub, iters=20, **kw):
  model = tm.load(name)
  q0, v0 = tm.initial_states(model, name, B, 0)
  phys = BatchedPhysics(model, batch=B, outputs=('xpos', 'xmat', 'subtree_com', 'sensordata', 'ncon', 'nefc', 'solver_niter'), full_final=True, **kw)
  phys.check_errors = False
  phys.data.qpos.copy_(torch.as_tensor(q0)); phys.data.qvel.copy_(torch.as_tensor(v0)); phys.forward()
  g = torch.Generator(device='cuda').manual_seed(0)
  for _ in range(5):
    phys.data.ctrl.uniform_(-1, 1, generator=g); phys.step(nsub)
  torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
  s.record()
  for _ in range(iters):
    phys.data.ctrl.uniform_(-1, 1, generator=g); phys.step(nsub)
  e.record(); torch.cuda.synchronize()
  ms = s.elapsed_time(e) / iters
  mx_con, mx_efc = 0, 0
  nit = []
  for _ in range(10):
    phys.data.ctrl.uniform_(-1, 1, generator=g); phys.step(nsub)
    mx_con = max(mx_con, int(phys.data.ncon.max())); mx_efc = max(mx_efc, int(phys.data.nefc.max())); nit.append(float(phys.data.solver_niter.float().mean()))
  print(f'   stats {name} {kw}: max ncon {mx_con} max nefc {mx_efc} mean nefc {float(phys.data.nefc.float().mean()):.1f} mean niter {sum(nit)/len(nit):.2f}', flush=True)
  print(f'BENCH {name}: B={B} nsub={nsub} {ms:.3f} ms/env-step-batch -> {B / ms * 1e3:.0f} env-steps/s  ws {phys.workspace_bytes()} B epb {phys.envs_per_block()} warn {phys.data.warning.sum(0).tolist()}', flush=True)

if __name__ == '__main__':
  print(torch.cuda.get_device_name(0), flush=True)
  for name, B, nc, nsub in (('cartpole', 16, 50, 1), ('pendulum_free', 8, 20, 2), ('cheetah', 32, 100, 1), ('humanoid', 32, 20, 5)):
    try:
      run(name, B, nc, nsub)
    except Exception as ex:
      import traceback; traceback.print_exc()
  for name, B, nsub, kw in (('cheetah', 4096, 1, {}), ('humanoid', 8192, 5, {}), ('humanoid', 8192, 5, dict(nconmax=24, njmax=64)), ('humanoid', 8192, 5, dict(nconmax=16, njmax=48))):
    try:
      bench(name, B, nsub, **kw)
    except Exception as ex:
      import traceback; traceback.print_exc()
