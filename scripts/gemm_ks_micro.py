"""micro-benchmark of acsr_gemm_batch (gemm_ks.cu) against linear_tok (d=64), the SIMT weight-gradient kernel and cuBLAS fp32."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ac_tsr_b200 as A

ops = A.ops
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device('cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=20, fl=True):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(n):
        if fl:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3


def fwd_case(T, K, N, name):
    X, W, b = torch.randn(T, K, device=dev), torch.randn(N, K, device=dev) * 0.1, torch.randn(N, device=dev)
    Y = torch.empty(T, N, device=dev)
    pr = [ops.gemm_problem(X, W, Y, T, N, K, bias=b)]
    t_new = timeit(lambda: ops.gemm_batch(pr))
    t_lib = timeit(lambda: torch.addmm(b, X, W.t(), out=Y))
    t_old = timeit(lambda: ops.linear_tok(X, T, K, W, N, Y, N, bias=b)) if K <= 256 else float('nan')
    fl = 2.0 * T * K * N
    print('%-28s T=%7d K=%4d N=%4d  gemm_ks %8.1f us (%6.1f TF)  linear_tok %8.1f us  cuBLAS fp32 %8.1f us' % (name, T, K, N, t_new, fl / t_new / 1e6, t_old, t_lib))


def wgrad_case(T, N, K, name):
    dY, X = torch.randn(T, N, device=dev), torch.randn(T, K, device=dev)
    dW, db = torch.zeros(N, K, device=dev), torch.zeros(N, device=dev)
    pr = [ops.wgrad_problem(dY, X, T, N, K, dW, db)]
    t_new = timeit(lambda: ops.gemm_batch(pr))
    t_old = timeit(lambda: ops.linear_wgrad(dY, X, dW, db))
    t_lib = timeit(lambda: dW.addmm_(dY.t(), X))
    print('%-28s T=%7d N=%4d K=%4d  gemm_ks %8.1f us  simt wgrad %8.1f us  cuBLAS fp32 %8.1f us' % (name, T, N, K, t_new, t_old, t_lib))


def layer_wgrads(T, d, I, L, name):
    mk = lambda *s: torch.randn(*s, device=dev)
    shapes = [(d, I), (I, d), (d, d), (d, d), (d, d), (L, d), (d, d), (d, d), (d, d)]
    prs, keep = [], []
    for (n, k) in shapes:
        dY, X, dW, db = mk(T, n), mk(T, k), torch.zeros(n, k, device=dev), torch.zeros(n, device=dev)
        keep.append((dY, X, dW, db))
        prs.append(ops.wgrad_problem(dY, X, T, n, k, dW, db))
    t_new = timeit(lambda: ops.gemm_batch(prs))
    def old():
        for dY, X, dW, db in keep:
            ops.linear_wgrad(dY, X, dW, db)
    t_old = timeit(old)
    print('%-28s T=%7d d=%d I=%d: 9 weight gradients of a layer, one launch %8.1f us ; 9 simt launches %8.1f us' % (name, T, d, I, t_new, t_old))


fwd_case(12800, 64, 64, 'C2 d->d')
fwd_case(12800, 64, 256, 'C2 ffn1')
fwd_case(12800, 256, 64, 'C2 ffn2')
fwd_case(25600, 64, 64, 'C2 2T d->d')
fwd_case(12800, 128, 128, 'C3v d->d')
fwd_case(12800, 128, 64, 'C3v ffn1')
fwd_case(102400, 64, 256, 'B=2048 ffn1')
fwd_case(409600, 256, 256, 'C5 d->d')
fwd_case(409600, 256, 1024, 'C5 ffn1')
fwd_case(409600, 1024, 256, 'C5 ffn2')
wgrad_case(12800, 64, 64, 'C2 dW d,d')
wgrad_case(12800, 256, 64, 'C2 dW1')
wgrad_case(12800, 64, 256, 'C2 dW2')
wgrad_case(409600, 256, 256, 'C5 dW d,d')
wgrad_case(409600, 1024, 256, 'C5 dW1')
layer_wgrads(12800, 64, 256, 50, 'C2 layer')
layer_wgrads(12800, 128, 64, 50, 'C3v layer')
