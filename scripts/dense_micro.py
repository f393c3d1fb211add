"""micro-benchmark of acsr_dense_fwd at the C2 shape (T = 12,800 tokens, inner 256, dropout 0.5 Philox)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
import ac_tsr_b200 as A
from ac_tsr_b200._lib import LIB
_p = A.ops._p
torch.manual_seed(0)
d, I, rows = 64, 256, int(sys.argv[1]) if len(sys.argv) > 1 else 12800
dev = 'cuda'
r = lambda *s: torch.randn(*s, device=dev)
Wo, W1, W2 = r(d, d) * 0.2, r(I, d) * 0.2, r(d, I) * 0.1
ops = torch.zeros(LIB.query('acsr_dense_prep_floats', I), device=dev)
st = torch.cuda.current_stream().cuda_stream
LIB.call('acsr_dense_prep', _p(Wo), _p(W1), _p(W2), d, I, _p(ops), st)
ctx, res = r(rows, d), r(rows, d)
z = lambda *s: torch.zeros(*s, device=dev)
S = dict(hz=z(rows, d), st_a=z(rows, 2), h=z(rows, d), z1=z(rows, I), a1=z(rows, I), z2=z(rows, d), st_f=z(rows, 2), out=z(rows, d))
one, zero = torch.ones(d, device=dev), torch.zeros(d, device=dev)
b1 = torch.zeros(I, device=dev)
rng = A.ops.DeviceRng(1, dev)
for p_drop in (0.5, 0.0):
    for act in (0, 1):
        def call():
            LIB.call('acsr_dense_fwd', _p(ctx), _p(res), rows, rows, d, I, act, _p(ops), _p(zero), _p(one), _p(zero), 1e-12, _p(b1),
                     _p(zero), _p(one), _p(zero), 1e-12, p_drop, None, None, rng.ptr, 3, 5, _p(S['hz']), _p(S['st_a']), _p(S['h']), _p(S['z1']),
                     _p(S['a1']), _p(S['z2']), _p(S['st_f']), _p(S['out']), 3, torch.cuda.current_stream().cuda_stream)
        call(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for _ in range(16):
                    call()
            g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); g.replay(); e1.record(); side.synchronize()
        print('rows %d p_drop %.1f act %d: %.1f us per launch' % (rows, p_drop, act, e0.elapsed_time(e1) * 1e3 / 32))
