"""ctypes binding of libacsr.so, generated from the prototypes in include/acsr.h.

There is NO fallback: if the library cannot be built/loaded, every op raises.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HEADER = os.path.join(ROOT, 'include', 'acsr.h')
LIB_PATH = os.path.join(HERE, 'csrc', 'libacsr.so')

_CTYPES = {
    'int': ctypes.c_int, 'float': ctypes.c_float, 'double': ctypes.c_double,
    'int64_t': ctypes.c_int64, 'uint32_t': ctypes.c_uint32, 'int32_t': ctypes.c_int32,
}


def parse_header(path=HEADER):
    """-> {name: (restype, [(argname, ctype)])} for every `acsr_*` prototype."""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    protos = {}
    for m in re.finditer(r'(const\s+char\s*\*|int)\s+(acsr_\w+)\s*\(([^)]*)\)\s*;', src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        restype = ctypes.c_char_p if 'char' in ret else ctypes.c_int
        argl = []
        args = ' '.join(args.split())
        if args and args != 'void':
            for a in args.split(','):
                a = a.strip()
                mm = re.match(r'(.*?)(\w+)$', a)
                typ, aname = mm.group(1).strip(), mm.group(2)
                if '*' in typ:
                    argl.append((aname, ctypes.c_void_p))
                else:
                    argl.append((aname, _CTYPES[typ.replace('const', '').strip()]))
        protos[name] = (restype, argl)
    return protos


class AcsrError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        self._dll = None
        self.protos = parse_header()
        self.timer = None            # KernelTimer while bench.py measures per-kernel device time

    def load(self):
        if self._dll is not None:
            return self._dll
        if not os.path.exists(LIB_PATH):
            raise AcsrError(
                'libacsr.so is not built (%s). Run `python -c "import __graft_entry__ as g; g.build()"` '
                'or `python ac-tsr_b200/build.py`. There is no CPU / PyTorch fallback for the hot path.' % LIB_PATH)
        dll = ctypes.CDLL(LIB_PATH)
        for name, (restype, argl) in self.protos.items():
            fn = getattr(dll, name)          # AttributeError if a declared symbol is missing
            fn.restype = restype
            fn.argtypes = [t for _, t in argl]
        if dll.acsr_version() != 1:
            raise AcsrError('libacsr.so ABI version mismatch')
        self._dll = dll
        return dll

    def call(self, name, *args):
        dll = self.load()
        if self.timer is not None:
            return self.timer.timed_call(self, dll, name, args)
        rc = getattr(dll, name)(*args)
        if rc != 0:
            msg = dll.acsr_last_error()
            raise AcsrError('%s failed (%d): %s' % (name, rc, msg.decode() if msg else ''))

    def query(self, name, *args):
        return getattr(self.load(), name)(*args)


class KernelTimer:
    """Brackets every C-ABI launch with CUDA events on the launching (current torch) stream.
    Used by bench.py for the per-kernel share / roofline numbers; never active in the timed step."""

    def __init__(self):
        self.events = []             # (name, start, stop)
        self.launches = 0

    def timed_call(self, lib, dll, name, args):
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(dll, name)(*args)
        e1.record()
        if rc != 0:
            msg = dll.acsr_last_error()
            raise AcsrError('%s failed (%d): %s' % (name, rc, msg.decode() if msg else ''))
        self.events.append((name, e0, e1))
        self.launches += 1

    def summary(self):
        """-> {name: (n_calls, total_ms)} after a device synchronize."""
        import torch
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1 in self.events:
            n, t = out.get(name, (0, 0.0))
            out[name] = (n + 1, t + e0.elapsed_time(e1))
        return out


LIB = _Lib()
