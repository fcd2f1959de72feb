"""The peer-memory primitives of the multi-GPU exchange (include/b200pir.h b200pir_peer_*): a buffer exported by one process is
mapped and written by ANOTHER process through a stream-ordered copy, as bench.py's ranks push their expanded queries into each
other's gather buffers.  One GPU suffices: CUDA IPC works between two processes on the same device."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import ctypes as C, sys
sys.path.insert(0, %r)
import torch
from sdk_b200._lib import LIB, check
handle = bytes.fromhex(sys.argv[1]); n = int(sys.argv[2])
ptr = C.c_void_p()
check(LIB.b200pir_peer_open(0, handle, C.byref(ptr)))
src = (torch.arange(n, dtype=torch.int64, device="cuda") * 2654435761 %% 251).to(torch.uint8)
s = torch.cuda.Stream()
check(LIB.b200pir_peer_copy_async(ptr.value + 16, src.data_ptr() + 16, n - 16, s.cuda_stream))     # an interior range, like a gather slot
s.synchronize()
check(LIB.b200pir_peer_close(0, ptr))
print("pushed")
""" % ROOT


def test_buffer_exported_by_one_process_is_written_by_another():
    import torch
    from sdk_b200._lib import LIB, check
    n = 1 << 20
    ptr, h = C.c_void_p(), C.create_string_buffer(64)
    check(LIB.b200pir_peer_alloc(0, n, C.byref(ptr), h))
    zero = torch.zeros(n, dtype=torch.uint8, device="cuda")
    check(LIB.b200pir_peer_copy_async(ptr.value, zero.data_ptr(), n, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    out = subprocess.run([sys.executable, "-c", CHILD, h.raw.hex(), str(n)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "pushed" in out.stdout, out.stderr[-2000:]
    back = torch.empty(n, dtype=torch.uint8, device="cuda")
    check(LIB.b200pir_peer_copy_async(back.data_ptr(), ptr.value, n, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    want = (np.arange(n, dtype=np.int64) * 2654435761 % 251).astype(np.uint8)
    want[:16] = 0
    assert np.array_equal(back.cpu().numpy(), want)
    check(LIB.b200pir_peer_free(0, ptr))
    with pytest.raises(Exception):
        check(LIB.b200pir_peer_open(0, b"\0" * 64, C.byref(C.c_void_p())))
