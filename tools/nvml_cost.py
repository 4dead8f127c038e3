import time, threading, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml as nv, torch
nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
for name, fn in (('clock', lambda: nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), ('reasons', lambda: nv.nvmlDeviceGetCurrentClocksEventReasons(h)),
                 ('maxclock', lambda: nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))):
  t = time.perf_counter(); [fn() for _ in range(20)]; print(name, 'ms per call %.3f' % ((time.perf_counter() - t) / 20 * 1e3))
exec(open('tools/step_timeline.py').read().split('# kernel time vs rollout age')[0].split("names = ['flush'")[0])
def run(n):
  s = torch.cuda.Event(True); e = torch.cuda.Event(True); torch.cuda.synchronize(); s.record()
  for _ in range(n):
    flush.fill_(0.0); a.uniform_(-1, 1, generator=g); env._task.before_step(a, phys); phys.step(5)
    r = env._task.get_reward(phys); o = env._task.get_observation(phys); pack(o, r)
  e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n
for _ in range(30): phys.step(5)
print('no sampler   %.3f ms/step' % run(40))
for period, what in ((0.2, 'clock+reasons'), (0.2, 'clock'), (1.0, 'clock+reasons')):
  stop = threading.Event()
  def poll():
    while not stop.is_set():
      nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
      if 'reasons' in what: nv.nvmlDeviceGetCurrentClocksEventReasons(h)
      stop.wait(period)
  th = threading.Thread(target=poll, daemon=True); th.start(); time.sleep(0.3)
  print('nvml %-14s every %.1fs: %.3f ms/step' % (what, period, run(40)))
  stop.set(); th.join()
