"""tcgen05.mma.kind::tf32 dispatch rate on B200 by N, operand source (SS / TS) and concurrent shared-memory traffic."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from ac_tsr_b200._lib import LIB_PATH
dll = ctypes.CDLL(LIB_PATH)
fn = dll.acsr_debug_umma_rate
fn.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
cyc = torch.zeros(148, dtype=torch.int64, device='cuda')
n = 4096
print('cycles per 128 x N x 8 TF32 MMA (148 CTAs, %d MMAs each); floor = N/2 cycles' % n)
for loaders in (0, 4):
    for mode, name in ((0, 'SS'), (1, 'TS (A in TMEM)')):
        row = []
        for N in (64, 128, 256):
            fn(N, mode, n, loaders, cyc.data_ptr(), 148, None)
            torch.cuda.synchronize()
            row.append('N=%d: %.1f' % (N, float(cyc.double().mean()) / n))
        print('  %-16s smem-traffic warps=%d   %s' % (name, loaders, '   '.join(row)))
