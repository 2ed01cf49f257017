"""Value model of the R subset: atomic vectors / matrices (numpy, column-major), lists, closures.

TEST INFRASTRUCTURE ONLY (part of oracle/).

R semantics kept: 1-based indexing, column-major recycling, NULL, named lists, names / dimnames / class
attributes, integer vs double.  Deliberate simplifications (none reachable with the arguments the fixtures use):
NA_real_ is NaN (is.na == is.nan for doubles), logical / integer NA are not represented, factors / S4 / environments
as values do not exist.
"""
import numpy as np


class RError(Exception):
    """stop()"""


class V:
    """Atomic vector or matrix.  a: ndarray, 1-D (vector) or 2-D Fortran-ordered (matrix);
    dtype float64 (double), int64 (integer), bool (logical) or object (character)."""
    __slots__ = ("a", "names", "dimnames", "attrs")

    def __init__(self, a, names=None, dimnames=None, attrs=None):
        if a.ndim == 2 and not a.flags.f_contiguous:
            a = np.asfortranarray(a)
        self.a = a
        self.names = names
        self.dimnames = dimnames
        self.attrs = attrs

    def __repr__(self):
        return f"V({self.a!r}, names={self.names}, dimnames={self.dimnames})"

    @property
    def kind(self):
        k = self.a.dtype.kind
        return {"f": "double", "i": "integer", "b": "logical", "O": "character", "U": "character"}[k]

    def flat(self):
        return self.a.reshape(-1, order="F") if self.a.ndim == 2 else self.a

    def __len__(self):
        return self.a.size


class RList:
    __slots__ = ("items", "names", "attrs")

    def __init__(self, items, names=None, attrs=None):
        self.items = list(items)
        self.names = list(names) if names is not None else None
        self.attrs = attrs

    def get(self, name):
        if self.names is None:
            return None
        for n, v in zip(self.names, self.items):
            if n == name:
                return v
        return None

    def __len__(self):
        return len(self.items)

    def __repr__(self):
        return f"RList({dict(zip(self.names or range(len(self.items)), self.items))})"


class Lang:
    """A call object as returned by match.call(): the callee and argument expressions."""
    __slots__ = ("exprs",)

    def __init__(self, exprs):
        self.exprs = list(exprs)


class Closure:
    __slots__ = ("params", "body", "env", "name")

    def __init__(self, params, body, env, name=None):
        self.params, self.body, self.env, self.name = params, body, env, name


class Builtin:
    __slots__ = ("fn", "name", "special")

    def __init__(self, fn, name, special=False):
        self.fn, self.name, self.special = fn, name, special  # special: receives (interp, env, arg exprs)


class Promise:
    __slots__ = ("expr", "env", "value", "forced")

    def __init__(self, expr, env):
        self.expr, self.env, self.forced, self.value = expr, env, False, None


class Env:
    __slots__ = ("vars", "parent", "call", "fn")

    def __init__(self, parent=None):
        self.vars = {}
        self.parent = parent
        self.call = None  # the ("call", ...) expression that created this frame
        self.fn = None


# ---------------------------------------------------------------------------------------- constructors / coercion
def dbl(x):
    return V(np.atleast_1d(np.asarray(x, dtype=np.float64)))


def intv(x):
    return V(np.atleast_1d(np.asarray(x, dtype=np.int64)))


def lgl(x):
    return V(np.atleast_1d(np.asarray(x, dtype=bool)))


def chr_(x):
    if isinstance(x, str):
        x = [x]
    a = np.empty(len(x), dtype=object)
    a[:] = list(x)
    return V(a)


def from_py(x):
    """numpy / python -> R value (matrices become Fortran-ordered)."""
    if x is None or isinstance(x, (V, RList, Closure, Builtin)):
        return x
    if isinstance(x, dict):
        return RList([from_py(v) for v in x.values()], list(x.keys()))
    if isinstance(x, str):
        return chr_(x)
    if isinstance(x, (bool, np.bool_)):
        return lgl(x)
    if isinstance(x, (int, np.integer)):
        return intv(x)
    if isinstance(x, (float, np.floating)):
        return dbl(x)
    a = np.asarray(x)
    if a.dtype.kind == "f":
        a = a.astype(np.float64, copy=False)
    elif a.dtype.kind in "iu":
        a = a.astype(np.int64)
    elif a.dtype.kind in "US":
        return chr_(list(a))
    return V(np.asfortranarray(a) if a.ndim == 2 else np.atleast_1d(a))


def to_py(v):
    """R value -> python: scalars stay length-1 arrays; lists become dicts (named) or lists."""
    if isinstance(v, V):
        return v.a
    if isinstance(v, RList):
        if v.names is not None and all(v.names):
            return {n: to_py(x) for n, x in zip(v.names, v.items)}
        return [to_py(x) for x in v.items]
    return v


def as_num(v):
    """numeric array view of an atomic vector (logical -> integer), for arithmetic."""
    a = v.a
    if a.dtype.kind == "b":
        return a.astype(np.int64)
    if a.dtype.kind == "O":
        raise RError("non-numeric argument to mathematical function")
    return a


def as_dbl(v):
    if v is None:
        return np.zeros(0)
    if isinstance(v, RList):
        return np.array([float(as_dbl(x)[0]) for x in v.items])
    a = v.a
    if a.dtype.kind == "O":
        return np.array([float(s) for s in a.reshape(-1, order="F")]).reshape(a.shape, order="F")
    return a.astype(np.float64, copy=False)


def scalar(v, what="argument"):
    if isinstance(v, V) and v.a.size >= 1:
        return v.flat()[0]
    raise RError(f"{what} is not a scalar value: {v!r}")


def truthy(v, what="condition"):
    if not isinstance(v, V) or v.a.size == 0:
        raise RError(f"{what}: argument is of length zero or not logical")
    x = v.flat()[0]
    if isinstance(x, (float, np.floating)) and x != x:
        raise RError(f"{what}: missing value where TRUE/FALSE needed")
    if v.a.size > 1:
        raise RError(f"{what}: the condition has length > 1")
    return bool(x)
