"""Host-side behaviour of the frame-scatter entry points (include/watsor_b200.h: wb_comm_*), no GPU needed: NCCL is
bound at run time, the rendezvous id comes out of the C-ABI, and loading it must not break a later `import torch`
(two libnccl.so.2 with different versions live in this image).  The collective itself is tested on two GPUs in
tests/test_gpu_scatter.py."""
import ctypes
import subprocess
import sys

from tests.conftest import ROOT
from watsor_b200 import _lib


def test_unique_id_comes_from_the_library():
    lib = _lib.load()
    a, b = (ctypes.c_uint8 * 128)(), (ctypes.c_uint8 * 128)()
    assert lib.wb_comm_unique_id(a) == 0, lib.wb_last_error()
    assert lib.wb_comm_unique_id(b) == 0
    assert bytes(a) != bytes(b) and any(bytes(a))
    assert lib.wb_comm_unique_id(None) != 0 and b'NULL' in lib.wb_last_error()


def test_calls_without_a_context_or_communicator_fail_loudly():
    lib = _lib.load()
    ident = (ctypes.c_uint8 * 128)()
    assert lib.wb_comm_init(None, 0, 1, ident) != 0
    assert lib.wb_scatter_frames(None, 0, None, None, 16, 0) != 0
    assert lib.wb_comm_destroy(None) == 0          # like wb_destroy(NULL)


def test_torch_still_imports_after_the_library_bound_nccl():
    code = ('import ctypes\n'
            'from watsor_b200 import _lib\n'
            'lib = _lib.load()\n'
            'buf = (ctypes.c_uint8 * 128)()\n'
            'assert lib.wb_comm_unique_id(buf) == 0\n'
            'import torch\n'
            'import torch.distributed\n'
            'assert lib.wb_comm_unique_id(buf) == 0\n'
            'print("ok")\n')
    out = subprocess.run([sys.executable, '-c', code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and 'ok' in out.stdout, out.stderr[-2000:]
