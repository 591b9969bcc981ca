"""ctypes binding of libibx.so, generated from ``include/ibx.h`` at import time.

The header is the single source of truth for the ABI: every prototype in it is parsed into
``argtypes``/``restype`` here, so a symbol the header declares but the library does not export fails
loudly at load (and in ``tests/test_abi.py``).  There is no fallback of any kind: if the shared
library is missing the import raises and tells the user to run ``__graft_entry__.build()``.
"""
import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "ibx.h")
LIBPATH = os.path.join(HERE, "libibx.so")


class IbxError(RuntimeError):
    """Raised for any non-zero status returned by the library (mirrors Julia's plain `error(...)`)."""


class Region(C.Structure):
    _fields_ = [("kind", C.c_int), ("h_is_f32", C.c_int), ("c", C.c_double * 3), ("a", C.c_double * 3),
                ("dfield", C.c_void_p), ("h", C.c_double)]


class SurfaceSpec(C.Structure):
    _fields_ = [("name", C.c_char_p), ("stl", C.c_void_p), ("h", C.c_double), ("h_is_f32", C.c_int),
                ("sphere_c", C.c_double * 3), ("sphere_r", C.c_double)]


class Fluid(C.Structure):
    _fields_ = [("R", C.c_float), ("gamma", C.c_float)]


class BCSpec(C.Structure):
    _fields_ = [("boundary", C.c_int), ("normal_flow", C.c_int), ("n_pinf", C.c_int), ("Pinf", C.c_float * 5)]


class Transport(C.Structure):
    _fields_ = [("mu_ref", C.c_float), ("T_ref", C.c_float), ("S", C.c_float), ("nk", C.c_int), ("k", C.c_float * 4)]


class WallParams(C.Structure):
    _fields_ = [("kappa", C.c_float), ("C", C.c_float), ("A", C.c_float), ("beta", C.c_float), ("beta_star", C.c_float),
                ("D", C.c_float), ("A_plus", C.c_float), ("omega", C.c_float), ("n_iter", C.c_int)]


_STRUCTS = {"ibx_transport": Transport, "ibx_wall_params": WallParams, "ibx_region": Region, "ibx_surface": SurfaceSpec, "ibx_fluid": Fluid, "ibx_bc_spec": BCSpec}
_OPAQUE = {"ibx_ctx", "ibx_stl", "ibx_dfield", "ibx_mesh", "ibx_domain", "ibx_accum", "void"}
_SCALARS = {"int": C.c_int, "int32_t": C.c_int32, "int64_t": C.c_int64, "float": C.c_float, "double": C.c_double,
            "ibx_array": C.c_int64, "char": C.c_char}


def _ctype(decl):
    """Map one C parameter declaration to a ctypes type."""
    decl = re.sub(r"/\*.*?\*/", "", decl).strip()
    if re.search(r"\[\s*\d*\s*\]$", decl):
        return C.c_void_p  # `char id[128]`: a raw buffer, not a NUL-terminated string
    stars = decl.count("*")
    words = [w for w in re.sub(r"[*]", " ", decl).split() if w != "const"]
    base = words[0]
    if base == "struct":
        base = words[1]
    if stars == 0:
        if base in _SCALARS:
            return _SCALARS[base]
        if base in _STRUCTS:
            return _STRUCTS[base]
        raise ValueError(f"cannot map by-value type in '{decl}'")
    if base == "char" and stars == 1:
        return C.c_char_p
    return C.c_void_p  # every other pointer (arrays, opaque handles, out-params) is passed as an address


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes])} for every prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    protos = {}
    for m in re.finditer(r"\b(int|const char\s*\*)\s+(ibx_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        argtypes = [] if args in ("", "void") else [_ctype(a) for a in args.split(",")]
        protos[name] = (C.c_char_p if "char" in ret else C.c_int, argtypes)
    return protos


PROTOTYPES = parse_header()


def _load():
    if not os.path.exists(LIBPATH):
        raise ImportError(f"{LIBPATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(libibx has no CPU fallback)")
    lib = C.CDLL(LIBPATH, mode=C.RTLD_GLOBAL)
    missing = []
    for name, (ret, argtypes) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = ret
        fn.argtypes = argtypes
    if missing and not os.environ.get("IBX_DEV_ALLOW_MISSING"):
        raise ImportError("libibx.so does not export symbols declared in include/ibx.h: " + ", ".join(missing))
    return lib


lib = _load()


def check(status):
    if status != 0:
        raise IbxError(f"[ibx status {status}] " + lib.ibx_last_error().decode("utf-8", "replace"))


def call(name, *args):
    check(getattr(lib, name)(*args))


def ptr(a):
    """Address of a numpy array (or None)."""
    return None if a is None else a.ctypes.data_as(C.c_void_p)
