import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


# B200MJ_EMULATE_GPU=1: run the `-m gpu` tests against the CPU emulation build of the kernels (tests/emu/gpu_shim.py).
# Off by default; on a GPU box nothing here changes anything.
EMULATE = os.environ.get('B200MJ_EMULATE_GPU') == '1'
DEV = 'cpu' if EMULATE else 'cuda'
if EMULATE:
  sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))
  import gpu_shim
  gpu_shim.install()


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu on the GPU box')


@pytest.fixture(scope='session')
def oracle_mod():
  from oracle import oracle
  oracle.build()
  return oracle
