"""tcgen05.mma.kind::tf32 dispatch rate on B200 by N, operand source (SS / TS), issuing threads and loop form."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from ac_tsr_b200._lib import LIB_PATH
dll = ctypes.CDLL(LIB_PATH)
fn = dll.acsr_debug_umma_rate
fn.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
cyc = torch.zeros(148, dtype=torch.int64, device='cuda')
n = 4096
print('cycles per 128 x N x 8 TF32 MMA of ONE issuing thread (148 CTAs, %d MMAs per issuer); tensor floor = N/2 cycles' % n)
for unrolled in (0, 1):
    for issuers in (1, 2):
        for mode, name in ((0, 'SS'), (1, 'TS (A in TMEM)')):
            row = []
            for N in (64, 128, 256):
                if issuers == 2 and N == 256:
                    continue
                fn(N, mode, n, 0, issuers, unrolled, cyc.data_ptr(), 148, None)
                torch.cuda.synchronize()
                per = float(cyc.double().mean()) / n
                row.append('N=%d: %.1f (%.1f per MMA overall)' % (N, per, per / issuers))
            print('  %-16s issuers=%d %-9s  %s' % (name, issuers, 'unrolled' if unrolled else 'rolled', '   '.join(row)))
