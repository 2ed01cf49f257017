"""Run the SHIPPED .Call shim (bindings/R/atlasqtl_b200_shim.c) without R -- TEST INFRASTRUCTURE.

The shim is compiled against the stand-in R API declarations of tests/r_stub together with the miniature runtime
tests/r_stub/rstub_runtime.c (tagged SEXP records, Rf_error -> longjmp, counted PROTECT stack, the routine table
R_init_atlasqtl registers) and linked against libatlasqtl_b200.so.  `RealShim` is then a `.Call` for the R evaluator of
oracle/rlite: it converts evaluator values to stub SEXPs WITHOUT copying double payloads (so in-place outputs land in the
evaluator's matrices, as under R), looks the symbol up in the registered table, checks the registered arity, calls the
wrapper, turns Rf_error() into an R error and reports a PROTECT imbalance as an error too.
"""
import ctypes
import os
import subprocess

import numpy as np

from oracle.rlite.values import RError, RList, V, chr_, dbl, intv

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STUB = os.path.join(HERE, "r_stub")
LIBDIR = os.path.join(ROOT, "atlasqtl_b200")
OUT = os.path.join(STUB, "_build", "libatlasqtl_rshim.so")
SHIM_SRC = os.path.join(ROOT, "bindings", "R", "atlasqtl_b200_shim.c")

NILSXP, LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP, EXTPTRSXP, RAWSXP = 0, 10, 13, 14, 16, 19, 22, 24


def build():
    srcs = [SHIM_SRC, os.path.join(STUB, "rstub_runtime.c")]
    deps = srcs + [os.path.join(STUB, "Rinternals.h"), os.path.join(ROOT, "include", "atlasqtl_b200.h"),
                   os.path.join(LIBDIR, "libatlasqtl_b200.so")]
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["gcc", "-O1", "-g", "-Wall", "-Wextra", "-Wno-cast-function-type", "-shared", "-fPIC", "-I", STUB,
           "-I", os.path.join(ROOT, "include"), *srcs, "-L", LIBDIR, "-latlasqtl_b200",
           "-Wl,-rpath,$ORIGIN/../../../atlasqtl_b200", "-o", OUT]   # relative: the tree may be copied elsewhere
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building the shim against the stub R runtime failed:\n" + res.stderr)
    return OUT


class StubPtr:
    """An R external pointer living in the stub runtime (the aq_ctx / aq_prep handle)."""

    def __init__(self, sexp):
        self.sexp = sexp


def load_library():
    try:
        return ctypes.CDLL(build())
    except OSError:           # a stale build from another location / toolchain: rebuild once
        if os.path.exists(OUT):
            os.remove(OUT)
        return ctypes.CDLL(build())


class RealShim:
    def __init__(self):
        self.lib = load_library()
        L = self.lib
        vp = ctypes.c_void_p
        L.rstub_wrap.restype = vp
        L.rstub_wrap.argtypes = [ctypes.c_int, vp, ctypes.c_ssize_t, ctypes.c_int, ctypes.c_int]
        L.rstub_nil.restype = vp
        L.Rf_allocVector.restype = vp
        L.Rf_allocVector.argtypes = [ctypes.c_uint, ctypes.c_ssize_t]
        L.SET_VECTOR_ELT.restype = vp
        L.SET_VECTOR_ELT.argtypes = [vp, ctypes.c_ssize_t, vp]
        L.VECTOR_ELT.restype = vp
        L.VECTOR_ELT.argtypes = [vp, ctypes.c_ssize_t]
        L.TYPEOF.argtypes = [vp]
        L.XLENGTH.restype = ctypes.c_ssize_t
        L.XLENGTH.argtypes = [vp]
        L.rstub_dim.argtypes = [vp, ctypes.c_int]
        L.rstub_data.restype = vp
        L.rstub_data.argtypes = [vp]
        L.rstub_names.restype = vp
        L.rstub_names.argtypes = [vp]
        L.rstub_string.restype = ctypes.c_char_p
        L.rstub_string.argtypes = [vp, ctypes.c_ssize_t]
        L.rstub_attr.restype = vp
        L.rstub_attr.argtypes = [vp, ctypes.c_char_p]
        L.rstub_finalize.argtypes = [vp]
        L.rstub_routine.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(vp),
                                    ctypes.POINTER(ctypes.c_int)]
        L.rstub_call.argtypes = [vp, ctypes.c_int, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.c_char_p, ctypes.c_int,
                                 ctypes.POINTER(ctypes.c_int)]
        L.R_MakeExternalPtr.restype = vp
        L.R_MakeExternalPtr.argtypes = [vp, vp, vp]
        L.R_init_atlasqtl.argtypes = [vp]
        L.R_init_atlasqtl(None)                     # what R does when it loads the package DLL
        self.routines = {}
        i = 0
        while True:
            name, fun, nargs = ctypes.c_char_p(), vp(), ctypes.c_int()
            if not L.rstub_routine(i, ctypes.byref(name), ctypes.byref(fun), ctypes.byref(nargs)):
                break
            self.routines[name.value.decode()] = (fun.value, nargs.value)
            i += 1
        self.calls = {}
        self._keep = []

    # ---- evaluator value -> SEXP
    def to_sexp(self, v):
        L = self.lib
        if v is None:
            return L.rstub_nil()
        if isinstance(v, StubPtr):
            return v.sexp
        if isinstance(v, RList):
            s = L.Rf_allocVector(VECSXP, len(v.items))
            for i, item in enumerate(v.items):
                L.SET_VECTOR_ELT(s, i, self.to_sexp(item))
            return s
        if not isinstance(v, V):
            raise RError(f".Call: cannot pass {type(v).__name__} to native code")
        a = v.a
        nrow, ncol = (a.shape if a.ndim == 2 else (-1, -1))
        if a.dtype == np.float64:
            if a.ndim == 2 and not a.flags.f_contiguous:
                raise RError(".Call: matrices must be column-major")
            buf = a if a.flags.c_contiguous or a.flags.f_contiguous else np.ascontiguousarray(a)
            self._keep.append(buf)
            return L.rstub_wrap(REALSXP, buf.ctypes.data, buf.size, nrow, ncol)   # no copy: in-place outputs reach `a`
        if a.dtype.kind in "ib":
            buf = np.asfortranarray(a.astype(np.int32))
            self._keep.append(buf)
            return L.rstub_wrap(LGLSXP if a.dtype.kind == "b" else INTSXP, buf.ctypes.data, buf.size, nrow, ncol)
        if a.dtype == np.uint8:
            buf = np.ascontiguousarray(a)
            self._keep.append(buf)
            return L.rstub_wrap(RAWSXP, buf.ctypes.data, buf.size, -1, -1)
        raise RError(f".Call: unsupported vector type {a.dtype}")

    # ---- SEXP -> evaluator value
    def from_sexp(self, s):
        L = self.lib
        t = L.TYPEOF(s)
        if t == NILSXP:
            return None
        if t == EXTPTRSXP:
            return StubPtr(s)
        n = L.XLENGTH(s)
        if t == VECSXP:
            items = [self.from_sexp(L.VECTOR_ELT(s, i)) for i in range(n)]
            nm = L.rstub_names(s)
            names = [L.rstub_string(nm, i).decode() for i in range(n)] if L.TYPEOF(nm) == STRSXP else None
            return RList(items, names)
        if t in (REALSXP, INTSXP, LGLSXP, RAWSXP):
            ct = {REALSXP: ctypes.c_double, INTSXP: ctypes.c_int, LGLSXP: ctypes.c_int, RAWSXP: ctypes.c_ubyte}[t]
            arr = np.ctypeslib.as_array(ctypes.cast(L.rstub_data(s), ctypes.POINTER(ct)), shape=(max(n, 1),))[:n].copy()
            nrow = L.rstub_dim(s, 0)
            if nrow >= 0:
                arr = arr.reshape((nrow, L.rstub_dim(s, 1)), order="F")
            if t == REALSXP:
                return V(arr if arr.ndim == 2 else np.atleast_1d(arr))
            if t == LGLSXP:
                return V(arr.astype(bool))
            return intv(arr.astype(np.int64)) if arr.ndim == 1 else V(arr.astype(np.int64))
        if t == STRSXP:
            return chr_([L.rstub_string(s, i).decode() for i in range(n)])
        raise RError(f".Call returned an object of type {t} the bridge does not handle")

    def attr(self, ptr, name):
        return self.from_sexp(self.lib.rstub_attr(ptr.sexp, name.encode()))

    def finalize(self, ptr):
        """What the garbage collector would do to an unreachable external pointer."""
        self.lib.rstub_finalize(ptr.sexp)

    # ---- .Call
    def __call__(self, it, pos, named):
        sym = pos[0].a[0]
        if sym not in self.routines:
            raise RError(f'.Call: "{sym}" is not a registered routine')        # R_useDynamicSymbols(dll, FALSE)
        fun, nargs = self.routines[sym]
        args = pos[1:]
        if len(args) != nargs:
            raise RError(f"Incorrect number of arguments ({len(args)}), expecting {nargs} for '{sym}'")
        self.calls[sym] = self.calls.get(sym, 0) + 1
        self._keep = []
        arr = (ctypes.c_void_p * max(nargs, 1))(*[self.to_sexp(a) for a in args])
        out, msg, imb = ctypes.c_void_p(), ctypes.create_string_buffer(1024), ctypes.c_int(0)
        rc = self.lib.rstub_call(fun, nargs, arr, ctypes.byref(out), msg, 1024, ctypes.byref(imb))
        if rc != 0:
            raise RError(msg.value.decode(errors="replace"))
        if imb.value != 0:
            raise RError(f"stack imbalance in '.Call' of {sym}: {imb.value}")
        return self.from_sexp(out.value)


def dbl_scalar(x):
    return dbl(float(x))
