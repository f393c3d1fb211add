"""Measure which shared-memory words a tcgen05.mma.kind::tf32 operand descriptor really addresses (debug tool).
The probed operand's region holds its own word index (split in two runs: low 10 bits / high bits, exact in TF32);
the other operand is an 8x8 identity in the known-good K-major no-swizzle layout, so D reads the probed operand back."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ac_tsr_b200 as A
dll = A.LIB.load()
fn = dll.acsr_debug_umma_probe
fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
fn.restype = ctypes.c_int

IMG = 128 * 1024            # bytes
ID_OFF = 0                  # identity operand at image offset 0 (K-major no swizzle: LBO = rows*16, SBO = 128)
PR_OFF = 32 * 1024          # probed operand region starts here (64 KB of unique ids)
PR_WORDS = 16 * 1024


def desc(start, lbo, sbo, layout=0):
    return ((start >> 4) & 0x3FFF) | (((lbo >> 4) & 0x3FFF) << 16) | (((sbo >> 4) & 0x3FFF) << 32) | (1 << 46) | (layout << 61)


def idesc(M, N, a_mn, b_mn):
    return (1 << 4) | (2 << 7) | (2 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def identity_image(rows):
    """K-major no-swizzle [rows x 8] operand with element (m,k) = 1 if m == k"""
    img = torch.zeros(IMG // 4)
    for k in range(8):
        m = k
        off = (k // 4) * (rows * 16) + m * 16 + (k % 4) * 4
        img[(ID_OFF + off) // 4] = 1.0
    return img


def run(probe_is_b, pdesc_fn, a_mn, b_mn, N=64):
    """-> word index map [n or m, k] of the probed operand (or None if the MMA was skipped)"""
    res = []
    for part in (0, 1):
        img = identity_image(128 if probe_is_b else N)
        ids = torch.arange(PR_WORDS)
        img[PR_OFF // 4: PR_OFF // 4 + PR_WORDS] = ((ids % 1024) if part == 0 else (ids // 1024)).float() + 1.0
        img = img.cuda()
        D = torch.zeros(128, N if N % 32 == 0 else 64).cuda()
        if probe_is_b:
            ad, bd = desc(ID_OFF, 128 * 16, 128), pdesc_fn(PR_OFF)
        else:
            ad, bd = pdesc_fn(PR_OFF), desc(ID_OFF, N * 16, 128)
        rc = fn(img.data_ptr(), IMG, ad, bd, idesc(128, N, a_mn, b_mn), D.shape[1], -777.0, D.data_ptr(), None)
        torch.cuda.synchronize()
        assert rc == 0
        res.append(D.cpu())
    lo, hi = res
    if bool((lo == -777.0).all()):
        return None
    if probe_is_b:      # D[m][n] = B(n, k=m), m < 8
        w = (lo[:8, :] - 1) + (hi[:8, :] - 1) * 1024            # [k, n]
        return w.t().contiguous()                               # [n, k]
    w = (lo[:, :8] - 1) + (hi[:, :8] - 1) * 1024                # [m, k]
    return w


def describe(name, w):
    if w is None:
        print('%-60s SKIPPED (sentinel intact)' % name); return
    if float(w.abs().max()) == 0 and float(w.min()) == 0:
        pass
    w = w.long()
    R = w.shape[0]
    print('%-60s rows %d' % (name, R))
    print('   byte offsets (row 0, k 0..7):', [int(x) * 4 for x in w[0]])
    print('   byte offsets (rows 0..9, k 0):', [int(x) * 4 for x in w[:10, 0]])
    print('   byte offsets (rows 30..35, k 0):', [int(x) * 4 for x in w[30:36, 0]])
    print('   row 1 k 0..7:', [int(x) * 4 for x in w[1]])
    if R > 64:
        print('   rows 64..67 k 0:', [int(x) * 4 for x in w[64:68, 0]])


print('=== B operand (N=64, K=8) ===')
describe('K-major none LBO=1024 SBO=128 (sanity)', run(True, lambda o: desc(o, 64 * 16, 128), 0, 0))
describe('K-major SW128 SBO=1024', run(True, lambda o: desc(o, 16, 1024, 2), 0, 0))
describe('MN-major none LBO=128 SBO=1024', run(True, lambda o: desc(o, 128, 1024), 0, 1))
describe('MN-major none LBO=1024 SBO=128', run(True, lambda o: desc(o, 1024, 128), 0, 1))
describe('MN-major SW128 LBO=8192 SBO=1024', run(True, lambda o: desc(o, 8192, 1024, 2), 0, 1))
describe('MN-major SW128 LBO=1024 SBO=8192', run(True, lambda o: desc(o, 1024, 8192, 2), 0, 1))
describe('MN-major SW128_BASE32B LBO=8192 SBO=1024', run(True, lambda o: desc(o, 8192, 1024, 1), 0, 1))
describe('MN-major SW64 LBO=8192 SBO=1024', run(True, lambda o: desc(o, 8192, 1024, 4), 0, 1))
describe('MN-major SW32 LBO=8192 SBO=1024', run(True, lambda o: desc(o, 8192, 1024, 6), 0, 1))
print('=== A operand (M=128, K=8), B identity N=64 ===')
describe('K-major none LBO=2048 SBO=128 (sanity)', run(False, lambda o: desc(o, 2048, 128), 0, 0))
describe('MN-major none LBO=128 SBO=2048', run(False, lambda o: desc(o, 128, 2048), 1, 0))
describe('MN-major SW128 LBO=16384 SBO=1024', run(False, lambda o: desc(o, 16384, 1024, 2), 1, 0))
describe('MN-major SW128_BASE32B LBO=16384 SBO=1024', run(False, lambda o: desc(o, 16384, 1024, 1), 1, 0))
