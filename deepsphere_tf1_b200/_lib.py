"""ctypes binding of libdsb200.so -- the C-ABI CUDA library of the ChebConv hot path.

The prototypes are read from include/dsb200.h (single source of truth).  There is NO CPU
fallback: if the library is missing or a kernel reports an error this module raises.
"""
import ctypes
import os
import re

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, '..', 'include', 'dsb200.h')
LIB_PATH = os.path.join(HERE, 'libdsb200.so')

_CTYPES = {
    'int': ctypes.c_int, 'long long': ctypes.c_longlong, 'float': ctypes.c_float, 'double': ctypes.c_double,
}


def parse_header(path=HEADER):
    """{name: (restype, [(ctype, argname)])} for every `dsb200_*` prototype in the header."""
    text = open(path).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    protos = {}
    for m in re.finditer(r'(const char\*|long long|int)\s+(dsb200_\w+)\s*\(([^)]*)\)\s*;', text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argl = []
        if args and args != 'void':
            for a in args.split(','):
                a = ' '.join(a.split())
                if '*' in a:
                    argl.append((ctypes.c_void_p, a.split('*')[-1].strip()))
                else:
                    ty, an = a.rsplit(' ', 1)
                    argl.append((_CTYPES[ty], an))
        protos[name] = ({'const char*': ctypes.c_char_p, 'long long': ctypes.c_longlong}.get(ret, ctypes.c_int), argl)
    return protos


PROTOS = parse_header()


class _Lib(object):
    def __init__(self):
        self._dll = None

    def load(self):
        if self._dll is not None:
            return self._dll
        if not os.path.exists(LIB_PATH):
            raise ImportError('libdsb200.so is not built (run `python -m deepsphere_tf1_b200.build`); '
                              'there is no CPU fallback for the ChebConv hot path')
        dll = ctypes.CDLL(LIB_PATH)
        for name, (ret, args) in PROTOS.items():
            fn = getattr(dll, name)                      # AttributeError if the symbol is missing
            fn.restype = ret
            fn.argtypes = [a[0] for a in args]
        if dll.dsb200_abi_version() != 2:
            raise ImportError('libdsb200.so ABI mismatch; rebuild')
        self._dll = dll
        return dll

    def last_error(self):
        return self.load().dsb200_last_error().decode()


LIB = _Lib()
launch_count = 0          # number of library calls issued (each is >= 1 kernel launch)


def kernel_launches():
    """CUDA kernels launched by libdsb200 in this process so far."""
    return int(LIB.load().dsb200_kernel_launches())


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        if not t.is_cuda:
            raise RuntimeError('libdsb200 needs CUDA tensors; there is no CPU fallback')
        if not t.is_contiguous():
            raise RuntimeError('libdsb200 needs contiguous tensors')
        return t.data_ptr()
    return t


def call(name, *args):
    """Call `dsb200_<name>(*args, stream)` on torch's current CUDA stream; raise on error."""
    global launch_count
    dll = LIB.load()
    fn = getattr(dll, 'dsb200_' + name)
    ptrs = [_ptr(a) for a in args]
    stream = torch.cuda.current_stream().cuda_stream
    rc = fn(*ptrs, stream)
    launch_count += 1
    if rc != 0:
        raise RuntimeError('dsb200_{} failed (code {}): {}'.format(name, rc, LIB.last_error()))
