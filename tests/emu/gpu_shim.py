"""Runs the `-m gpu` tests on a machine WITHOUT a GPU: B200MJ_EMULATE_GPU=1 python -m pytest tests -m gpu.

Test infrastructure only, off unless that variable is set. It points `dm_control_b200.lib` at the CPU emulation build
of the kernel source (cuda_emu.h), makes `BatchedPhysics` allocate its tensors on the host, and stubs the handful of
`torch.cuda` calls the facade makes. What passes this way has exercised the Python facade, the C ABI and the kernels'
logic — not the device: the same tests on a B200 remain the parity tests proper.
"""
import contextlib
import types


def install():
  import torch
  import b200mj_emu as emu
  from dm_control_b200 import lib as blib, physics
  blib.SO_PATH = emu.build()
  blib._lib = None
  stream = types.SimpleNamespace(cuda_stream=0, synchronize=lambda: None, wait_stream=lambda s: None)
  torch.cuda.is_available = lambda: True
  torch.cuda.current_device = lambda: 0
  torch.cuda.device = lambda d=None: contextlib.nullcontext()
  torch.cuda.current_stream = lambda d=None: stream
  torch.cuda.synchronize = lambda *a, **k: None
  torch.Tensor.pin_memory = lambda self, *a, **k: self
  orig_init = physics.BatchedPhysics.__init__

  def init(self, model, batch=1, device=None, *a, **k):
    orig_init(self, model, batch, 'cpu', *a, **k)
  physics.BatchedPhysics.__init__ = init
