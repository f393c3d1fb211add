import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
import ac_tsr_b200 as A
from ac_tsr_b200._lib import LIB
_p = A.ops._p
torch.manual_seed(0)
d, I, rows = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = 'cuda'
Wo = torch.randn(d, d, device=dev) * 0.2
W1 = torch.randn(I, d, device=dev) * 0.2
W2 = torch.randn(d, I, device=dev) * 0.1
ops = torch.zeros(LIB.query('acsr_dense_prep_floats', I), device=dev)
st = torch.cuda.current_stream().cuda_stream
LIB.call('acsr_dense_prep', _p(Wo), _p(W1), _p(W2), d, I, _p(ops), st)
torch.cuda.synchronize()
blk = ops.view(-1, 2, 16, 64, 4)          # [block][hi|lo][kc][n][4]
rec = (blk[:, 0] + blk[:, 1]).permute(0, 2, 1, 3).reshape(-1, 64, 64)      # [block][n][k]
print('prep Wo ok', torch.equal(rec[0], Wo), 'W1', torch.equal(rec[1], W1[:64]), 'W2', torch.equal(rec[1 + I // 64], W2[:, :64]))
ctx = torch.randn(rows, d, device=dev)
res = torch.zeros(rows, d, device=dev)
z = lambda *s: torch.zeros(*s, device=dev)
S = dict(hz=z(rows, d), st_a=z(rows, 2), h=z(rows, d), z1=z(rows, I), a1=z(rows, I), z2=z(rows, d), st_f=z(rows, 2), out=z(rows, d))
one, zero = torch.ones(d, device=dev), torch.zeros(d, device=dev)
LIB.call('acsr_dense_fwd', _p(ctx), _p(res), rows, rows, d, I, 1, _p(ops), _p(zero), _p(one), _p(zero), 1e-12, _p(torch.zeros(I, device=dev)),
         _p(zero), _p(one), _p(zero), 1e-12, 0.0, None, None, None, 0, 0, _p(S['hz']), _p(S['st_a']), _p(S['h']), _p(S['z1']), _p(S['a1']),
         _p(S['z2']), _p(S['st_f']), _p(S['out']), 3, st)
torch.cuda.synchronize()
ref = ctx.double() @ Wo.double().t()
got = S['hz'].double()
print('hz max err', float((got - ref).abs().max()), 'scale', float(ref.abs().max()))
# which structure? compare with transposed / permuted hypotheses
print('hz vs ctx@Wo (no transpose)', float((got - ctx.double() @ Wo.double()).abs().max()))
print('row 0 got', got[0, :6].tolist()); print('row 0 ref', ref[0, :6].tolist())
print('col sums got/ref', float(got.sum()), float(ref.sum()))
href = torch.nn.functional.layer_norm(ref, (d,))
print('h err', float((S['h'].double() - href).abs().max()))

z1ref = href @ W1.double().t()
a1ref = torch.relu(z1ref)
z2ref = a1ref @ W2.double().t()
for name, ref_, got_ in (('hz', ref, S['hz']), ('h', href, S['h']), ('z1', z1ref, S['z1']), ('a1', a1ref, S['a1']), ('z2', z2ref, S['z2'])):
    e = (got_.double() - ref_).abs()
    per_tile = [float(e[i:i + 128].max()) for i in range(0, rows, 128)]
    print(name, 'max err per tile', ['%.2e' % v for v in per_tile])
    if name in ('z1', 'a1'):
        print('   per 64-col chunk', ['%.2e' % float(e[:, c:c + 64].max()) for c in range(0, I, 64)])
