"""ctypes binding of libvadc.so — the C-ABI boundary (include/vadc.h).

The prototypes are parsed from the header itself so the header stays the
single source of truth.  There is NO fallback: if the library is missing or a
call fails, a RuntimeError is raised."""
import ctypes
import os
import re
import threading

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "vadc.h")
LIB_PATH = os.environ.get("VADC_LIB_PATH") or os.path.join(HERE, "libvadc.so")   # override: kernel-variant experiments

IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05 = 0, 1, 2
LOSS_L1_MEAN, LOSS_MSE_MEAN, LOSS_E4_NORM = 0, 1, 2

_CTYPES = {
    "int": ctypes.c_int, "int64_t": ctypes.c_int64, "float": ctypes.c_float,
    "size_t": ctypes.c_size_t, "unsigned long long": ctypes.c_ulonglong,
    "double": ctypes.c_double, "void": None,
}


def _ctype(decl):
    decl = decl.replace("const ", "").strip()
    if decl.endswith("*"):
        return ctypes.c_char_p if decl[:-1].strip() == "char" else ctypes.c_void_p
    return _CTYPES[decl]


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes])} for every function declared in vadc.h"""
    with open(path) as fh:
        src = fh.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = "\n".join(ln for ln in src.splitlines() if not ln.lstrip().startswith("#"))
    src = src.replace('extern "C" {', "")
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(vadc_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)(\w+)$", a)          # strip the parameter name
                argtypes.append(_ctype(mm.group(1).strip()))
        protos[name] = (_ctype(ret), argtypes)
    return protos


_lib = None
_tls = threading.local()          # devices of the tensors whose pointers were taken for the call being assembled


class _CurrentStream:
    """placeholder returned by ``stream()``: resolved to the current stream of the DEVICE THE TENSOR ARGUMENTS LIVE
    ON when the library function is called (libvadc launches on whatever context cudaGetDevice() reports)"""


_CUR_STREAM = _CurrentStream()


def _guarded(fn, name):
    """Every kernel-launching entry point runs with the tensors' own device current and on that device's current
    stream: a module on cuda:1 without set_device, an autograd thread of another device, DataParallel replicas.
    All tensor arguments of one call must share one device."""
    def call(*args):
        devs = getattr(_tls, "devs", None) or ()
        _tls.devs = []
        d0 = devs[0] if devs else None
        if d0 is not None and any(d != d0 for d in devs):
            raise RuntimeError(f"{name}: tensor arguments live on different CUDA devices {sorted(set(devs))}")
        if d0 is None and not any(a is _CUR_STREAM for a in args):
            return fn(*args)                                  # *_workspace_bytes, error strings, counters
        dev = d0 if d0 is not None else torch.cuda.current_device()
        with torch.cuda.device(dev):
            args = tuple(ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream) if a is _CUR_STREAM else a
                         for a in args)
            return fn(*args)
    call.__name__ = name
    call.raw = fn
    return call


class _Lib:
    """attribute access returns the device-guarded callables; ``_cdll`` is the raw ctypes handle"""

    def __init__(self, cdll, names):
        self._cdll = cdll
        for name in names:
            setattr(self, name, _guarded(getattr(cdll, name), name))


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
            "(nvcc, sm_100a). videoad_b200 has no CPU or PyTorch fallback.")
    l = ctypes.CDLL(LIB_PATH)
    protos = parse_header()
    for name, (res, args) in protos.items():
        fn = getattr(l, name)            # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = _Lib(l, protos.keys())
    return _lib


def check(rc, what):
    if rc != 0:
        l = lib()
        msg = l.vadc_error_string(rc).decode()
        if rc == -5:
            msg += ": " + l.vadc_last_cuda_error().decode()
        raise RuntimeError(f"{what} failed: {msg} (code {rc})")


def require_cuda(*tensors):
    _tls.devs = []                       # a new op starts: forget pointers taken by a call that never happened
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "videoad_b200 runs on sm_100a CUDA tensors only (no CPU fallback); got a "
                f"{t.device} tensor")


def ptr(t):
    """device pointer of ``t`` for the call being assembled; remembers the device for the guard in ``_guarded``"""
    if t is None:
        return None
    if t.is_cuda:
        devs = getattr(_tls, "devs", None)
        if devs is None:
            devs = _tls.devs = []
        devs.append(t.device.index)
    return ctypes.c_void_p(t.data_ptr())


def stream():
    """the current stream of the device the call's tensors live on (resolved at call time)"""
    return _CUR_STREAM


def f32c(t):
    """fp32 + contiguous + 16-byte aligned base (no copy when already so).  A contiguous view with an odd storage
    offset (a sliced batch) is cloned: the kernels use 128-bit accesses and TMA."""
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone(memory_format=torch.contiguous_format)
    return t


_ws_cache = {}


def workspace(nbytes, device):
    """stream-ordered scratch buffer, grown on demand and reused per (device, stream)"""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def launch_count():
    return int(lib().vadc_launch_count())
