"""ctypes binding of libxsb200.so (the C ABI declared in include/xsb200.h).

Prototypes are parsed from the header, so the header is the single source of truth for the
boundary. There is NO fallback: if the shared library is missing or a call fails, the op raises.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
HEADER = os.path.join(ROOT, "include", "xsb200.h")
LIB_PATH = os.path.join(_HERE, "libxsb200.so")

# fused entry points that answer XSB_ERR_UNSUPPORTED (3) for shapes outside their envelope instead of failing
_MAY_DECLINE = {"xsb_bn_relu_maxpool_fwd", "xsb_bn_bwd_pool_dxsum", "xsb_bn_add_relu"}

MATH_FP32 = 0
MATH_TF32 = 1

_CTYPES = {
    "int": ctypes.c_int,
    "int64_t": ctypes.c_int64,
    "float": ctypes.c_float,
    "void": None,
    "const char*": ctypes.c_char_p,
}


def _ctype_of(decl: str):
    decl = decl.strip()
    if decl.endswith("*") or "*" in decl:
        return ctypes.c_void_p
    return _CTYPES[decl]


def parse_header(path: str = HEADER):
    """-> {name: (restype, [argtypes])} for every `xsb_*` function declared in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"#[^\n]*", " ", src)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|int64_t|int)\s+(xsb_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        ret = "const char*" if "char" in ret else ret
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    argtypes.append(_CTYPES[a.rsplit(" ", 1)[0].replace("const ", "").strip()])
        protos[name] = (_CTYPES[ret], argtypes)
    return protos


class _Lib:
    def __init__(self):
        self._dll = None
        self.protos = None

    def load(self):
        if self._dll is not None:
            return self._dll
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). xshinnosuke_b200 has no CPU or PyTorch fallback.")
        dll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (restype, argtypes) in self.protos.items():
            fn = getattr(dll, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        self._dll = dll
        return dll

    def __getattr__(self, name):
        dll = self.load()
        fn = getattr(dll, name)
        if self.protos[name][0] is ctypes.c_int and name not in ("xsb_version", "xsb_device_sm_count"):
            def checked(*args, _fn=fn, _name=name):
                rc = _fn(*args)
                if rc == 3 and _name in _MAY_DECLINE:
                    return rc           # XSB_ERR_UNSUPPORTED, nothing launched: the caller runs the unfused kernels
                if rc != 0:
                    raise RuntimeError(f"{_name} failed (code {rc}): {dll.xsb_last_error().decode()}")
                return rc
            setattr(self, name, checked)
            return checked
        setattr(self, name, fn)
        return fn


lib = _Lib()
