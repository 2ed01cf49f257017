"""Built-in functions of the R subset: the parts of base / stats / gsl the reference's R files call.

TEST INFRASTRUCTURE ONLY (part of oracle/).  Written from the R documentation (?Arithmetic, ?Extract, ?sweep,
?scale, ?colSums, ?all.equal, ?order, ?duplicated, ?uniroot ...), not from the reference package.

Numerics: reductions (sum, colSums, rowSums, mean, cumsum) accumulate in extended precision like R's LDOUBLE
accumulators; crossprod / %*% go to BLAS as in R; the special functions are SciPy's (pnorm(log.p) -> log_ndtr,
digamma, lgamma -> gammaln, gsl::expint_E1 -> exp1, gsl::gamma_inc(a, x) -> gamma(a) gammaincc(a, x) for a > 0), each
pinned against 40-digit mpmath evaluations in tests/test_special_functions.py.
"""
import math
import os

import numpy as np
from scipy import special as sp

from . import parser as P
from .values import (Builtin, Closure, Env, Lang, RError, RList, V, as_dbl, as_num, chr_, dbl, intv, lgl, scalar,
                     truthy)

LD = np.longdouble


# ======================================================================================== arithmetic with recycling
def _attrs_from(x, y, shape):
    names = dimnames = None
    if len(shape) == 2:
        for v in (x, y):
            if v.a.ndim == 2 and v.dimnames is not None:
                dimnames = v.dimnames
                break
    else:
        for v in (x, y):
            if v.names is not None and v.a.size == shape[0]:
                names = v.names
                break
    return names, dimnames


def recycle2(x, y):
    a, b = x.a, y.a
    if a.shape == b.shape:
        return a, b, a.shape
    if a.ndim == 2 and b.ndim == 2:
        raise RError("non-conformable arrays")
    if a.ndim == 2 or b.ndim == 2:
        m, v, swap = (a, b, False) if a.ndim == 2 else (b, a, True)
        if v.size == 1:
            vv = v.reshape(1, 1)
        elif v.size == 0:
            raise RError("zero-length vector against a matrix")
        else:
            if v.size > m.size:
                raise RError("dims do not match the length of object")
            vv = np.resize(v, m.size).reshape(m.shape, order="F")
        return (vv, m, m.shape) if swap else (m, vv, m.shape)
    if a.size == 0 or b.size == 0:
        return a[:0], b[:0], (0,)
    n = max(a.size, b.size)
    return (a if a.size == n else np.resize(a, n)), (b if b.size == n else np.resize(b, n)), (n,)


def _mk(res, x, y, shape):
    names, dimnames = _attrs_from(x, y, shape)
    if res.shape != tuple(shape):
        res = np.broadcast_to(res, shape).copy(order="F")
    return V(res, names, dimnames)


def arith(op):
    def f(it, pos, named):
        x, y = pos
        if x is None or y is None:
            return dbl(np.zeros(0))
        if not isinstance(x, V) or not isinstance(y, V):
            raise RError(f"non-numeric argument to binary operator {op}")
        a, b, shape = recycle2(V(as_num(x), x.names, x.dimnames), V(as_num(y), y.names, y.dimnames))
        both_int = a.dtype.kind == "i" and b.dtype.kind == "i"
        with np.errstate(all="ignore"):
            if op == "+":
                r = a + b
            elif op == "-":
                r = a - b
            elif op == "*":
                r = a * b
            elif op == "/":
                r = a.astype(np.float64) / b
            elif op == "^":
                r = np.power(a.astype(np.float64), b.astype(np.float64))
                # R: 1 ^ y and x ^ 0 are 1 even for NA/NaN -- numpy agrees
            elif op == "%%":
                r = np.mod(a, b) if both_int else np.mod(a.astype(np.float64), b)
            elif op == "%/%":
                r = np.floor_divide(a, b) if both_int else np.floor(a.astype(np.float64) / b)
            else:
                raise RError(op)
        return _mk(np.asarray(r), x, y, shape)
    return f


def compare(op):
    fn = {"==": np.equal, "!=": np.not_equal, "<": np.less, ">": np.greater, "<=": np.less_equal,
          ">=": np.greater_equal}[op]

    def f(it, pos, named):
        x, y = pos
        if x is None or y is None:
            return lgl(np.zeros(0, bool))
        a, b, shape = recycle2(x, y)
        if a.dtype.kind == "O" or b.dtype.kind == "O":
            a = np.array([str(s) if not isinstance(s, str) else s for s in a.reshape(-1)], dtype=object).reshape(a.shape)
            b = np.array([str(s) if not isinstance(s, str) else s for s in b.reshape(-1)], dtype=object).reshape(b.shape)
        with np.errstate(invalid="ignore"):
            r = fn(a, b)
        return _mk(np.asarray(r, dtype=bool), x, y, shape)
    return f


def logic(op):
    def f(it, pos, named):
        x, y = pos
        a, b, shape = recycle2(V(x.a.astype(bool), x.names, x.dimnames), V(y.a.astype(bool), y.names, y.dimnames))
        return _mk(np.logical_and(a, b) if op == "&" else np.logical_or(a, b), x, y, shape)
    return f


def unary_minus(it, pos, named):
    x = pos[0]
    return V(-as_num(x), x.names, x.dimnames)


def unary_plus(it, pos, named):
    return pos[0]


def unary_not(it, pos, named):
    x = pos[0]
    return V(~x.a.astype(bool), x.names, x.dimnames)


def math1(fn, keep_int=False):
    def f(it, pos, named):
        x = pos[0]
        if x is None:
            raise RError("non-numeric argument to mathematical function")
        with np.errstate(all="ignore"):
            r = fn(as_num(x) if keep_int else as_dbl(x))
        return V(np.asarray(r), x.names, x.dimnames)
    return f


# ======================================================================================== indexing
def _resolve(idx, n, names, what="subscript"):
    """R subscript -> 0-based integer positions along an extent of length n (Ellipsis = everything)."""
    if idx is Ellipsis:
        return np.arange(n)
    if idx is None:
        return np.zeros(0, dtype=np.int64)
    a = idx.flat()
    k = a.dtype.kind
    if k == "b":
        if a.size > n:
            raise RError(f"(subscript) logical subscript too long ({a.size} > {n})")
        if a.size != n:
            a = np.resize(a, n)
        return np.flatnonzero(a)
    if k == "O":
        if names is None:
            raise RError(f"{what} out of bounds (no names)")
        lut = {}
        for i, nm in enumerate(names):
            lut.setdefault(nm, i)
        try:
            return np.array([lut[s] for s in a], dtype=np.int64)
        except KeyError as e:
            raise RError(f"{what} out of bounds: {e}")
    if k == "f":
        if np.isnan(a).any():
            raise RError("NA subscripts are not supported")
        a = np.trunc(a).astype(np.int64)
    if a.size and (a < 0).any():
        if (a > 0).any():
            raise RError("can't mix positive and negative subscripts")
        keep = np.ones(n, dtype=bool)
        drop = -a[a != 0] - 1
        keep[drop[drop < n]] = False
        return np.flatnonzero(keep)
    a = a[a != 0] - 1
    return a


def _na_of(dtype):
    return None if dtype.kind == "O" else np.nan


def index(it, obj, pos, named, double):
    drop = True
    if "drop" in named:
        drop = truthy(named["drop"], "drop")
    if "exact" in named:
        pass
    if obj is None:
        return None
    if isinstance(obj, Lang):
        keep = _resolve(pos[0], len(obj.exprs), None)
        return Lang([obj.exprs[i] for i in keep])
    if double:
        if isinstance(obj, RList):
            i = pos[0]
            if i.a.dtype.kind == "O":
                return obj.get(i.a[0])
            k = int(scalar(i)) - 1
            if k < 0 or k >= len(obj.items):
                raise RError("subscript out of bounds")
            return obj.items[k]
        if len(pos) == 1:
            k = _resolve(pos[0], obj.a.size, obj.names)
            if k.size != 1 or k[0] >= obj.a.size:
                raise RError("subscript out of bounds")
            return V(obj.flat()[k])
        pos = list(pos)
    if isinstance(obj, RList):
        keep = _resolve(pos[0], len(obj.items), obj.names)
        return RList([obj.items[i] for i in keep], None if obj.names is None else [obj.names[i] for i in keep],
                     None)
    a = obj.a
    if len(pos) == 1:
        i = pos[0]
        if i is Ellipsis:
            return obj
        if a.ndim == 2 and isinstance(i, V) and i.a.ndim == 2 and i.a.dtype.kind == "b":
            return V(obj.flat()[i.flat()])
        flat = obj.flat()
        names = obj.names if a.ndim == 1 else None
        k = _resolve(i, flat.size, names)
        oob = k >= flat.size
        if oob.any():
            res = np.empty(k.size, dtype=np.float64 if flat.dtype.kind in "ib" else flat.dtype)
            res[~oob] = flat[k[~oob]]
            res[oob] = _na_of(res.dtype)
            nm = None if names is None else [names[j] if j < flat.size else None for j in k]
            return V(res, nm)
        return V(flat[k], None if names is None else [names[j] for j in k])
    if len(pos) != 2 or a.ndim != 2:
        raise RError("incorrect number of dimensions")
    dn = obj.dimnames or [None, None]
    r = _resolve(pos[0], a.shape[0], dn[0], "row subscript")
    c = _resolve(pos[1], a.shape[1], dn[1], "column subscript")
    if (r >= a.shape[0]).any() or (c >= a.shape[1]).any():
        raise RError("subscript out of bounds")
    sub = a[np.ix_(r, c)]
    rn = None if dn[0] is None else [dn[0][j] for j in r]
    cn = None if dn[1] is None else [dn[1][j] for j in c]
    if drop and (sub.shape[0] == 1 or sub.shape[1] == 1):
        if sub.shape[0] == 1 and sub.shape[1] == 1:
            return V(sub.reshape(-1))
        return V(sub.reshape(-1, order="F"), cn if sub.shape[0] == 1 else rn)
    return V(np.asfortranarray(sub), None, None if rn is None and cn is None else [rn, cn])


def _promote(dst, src):
    """array dtype able to hold both (logical < integer < double < character)."""
    order = {"b": 0, "i": 1, "f": 2, "O": 3}
    kd, ks = dst.dtype.kind, src.dtype.kind
    if order[ks] > order[kd]:
        if ks == "O":
            out = np.empty(dst.shape, dtype=object)
            out[...] = dst
            return out
        return dst.astype(src.dtype)
    return dst.copy(order="F") if dst.ndim == 2 else dst.copy()


def _fill_values(it, val, count):
    v = val.flat() if isinstance(val, V) else np.asarray(val)
    if count == 0:
        return v[:0]
    if v.size == 0:
        raise RError("replacement has length zero")
    if v.size != count:
        if count % v.size != 0 or v.size > count:
            it.warn("number of items to replace is not a multiple of replacement length")
        v = np.resize(v, count)
    return v


def index_assign(it, obj, pos, named, double, val):
    if isinstance(obj, RList) or (obj is None and (double and not isinstance(val, V))) or \
            (obj is None and isinstance(val, RList)):
        lst = RList(obj.items, obj.names, obj.attrs) if obj is not None else RList([], None)
        i = pos[0]
        if not double:
            keep = _resolve(i, len(lst.items), lst.names)
            vals = val.items if isinstance(val, RList) else [val] * len(keep)
            for j, k in enumerate(keep):
                lst.items[k] = vals[j % len(vals)]
            return lst
        if i.a.dtype.kind == "O":
            return dollar_assign(lst, i.a[0], val)
        k = int(scalar(i)) - 1
        while len(lst.items) <= k:
            lst.items.append(None)
            if lst.names is not None:
                lst.names.append("")
        lst.items[k] = val
        return lst
    if not isinstance(val, V):
        raise RError("replacement value must be an atomic vector")
    if obj is None:   # NULL[i] <- value creates the vector (NA where nothing is assigned)
        n0 = pos[0].a.size if len(pos) == 1 and isinstance(pos[0], V) and pos[0].a.dtype.kind == "b" else 0
        init = np.empty(n0, dtype=val.a.dtype if val.a.dtype.kind in "fO" else np.float64)
        init[:] = _na_of(init.dtype)
        obj = V(init)
    a = _promote(obj.a, val.a)
    if len(pos) == 1:
        i = pos[0]
        flat = a.reshape(-1, order="F") if a.ndim == 2 else a
        names = obj.names
        if i is Ellipsis:
            k = np.arange(flat.size)
        elif isinstance(i, V) and i.a.dtype.kind == "b" and i.a.size > flat.size:
            raise RError("logical subscript too long")
        else:
            k = _resolve(i, flat.size, names if a.ndim == 1 else None)
        if k.size and k.max() >= flat.size:
            if a.ndim == 2:
                raise RError("subscript out of bounds")
            new = np.empty(k.max() + 1, dtype=flat.dtype)
            new[:flat.size] = flat
            new[flat.size:] = _na_of(flat.dtype) if flat.dtype.kind in "fO" else 0
            if names is not None:
                names = list(names) + [""] * (new.size - flat.size)
            flat = new
        flat = flat.copy() if a.ndim == 2 else flat
        flat[k] = _fill_values(it, val, k.size)
        if a.ndim == 2:
            return V(flat.reshape(a.shape, order="F"), None, obj.dimnames, obj.attrs)
        return V(flat, names, None, obj.attrs)
    if len(pos) != 2 or a.ndim != 2:
        raise RError("incorrect number of subscripts")
    dn = obj.dimnames or [None, None]
    r = _resolve(pos[0], a.shape[0], dn[0])
    c = _resolve(pos[1], a.shape[1], dn[1])
    vals = _fill_values(it, val, r.size * c.size)
    a[np.ix_(r, c)] = vals.reshape((r.size, c.size), order="F")
    return V(a, None, obj.dimnames, obj.attrs)


def dollar(obj, name):
    if obj is None:
        return None
    if not isinstance(obj, RList):
        raise RError("$ operator is invalid for atomic vectors")
    v = obj.get(name)
    if v is None and obj.names is not None:  # partial matching, as `$` does
        c = [n for n in obj.names if n and n.startswith(name)]
        if len(c) == 1:
            return obj.get(c[0])
    return v


def dollar_assign(obj, name, val):
    lst = RList(obj.items, obj.names, obj.attrs) if obj is not None else RList([], [])
    if not isinstance(lst, RList):
        raise RError("invalid type for $ assignment")
    if lst.names is None:
        lst.names = [""] * len(lst.items)
    if name in lst.names:
        k = lst.names.index(name)
        if val is None:
            del lst.items[k], lst.names[k]
        else:
            lst.items[k] = val
    elif val is not None:
        lst.items.append(val)
        lst.names.append(name)
    return lst


# ======================================================================================== helpers
def _arg(pos, named, k, name, default=None):
    if name in named:
        return named[name]
    if k is not None and k < len(pos):
        return pos[k]
    return default


def _flag(v, default):
    return default if v is None else truthy(v)


def _ldsum(a, axis=None):
    return np.sum(a.astype(LD), axis=axis).astype(np.float64)


def _strs(v):
    if v is None:
        return []
    if isinstance(v, RList):
        return [_fmt1(x) for x in v.items]
    return [_fmt_elem(x, 15) for x in v.flat()]


def _fmt_elem(x, digits=7):
    if x is None:
        return "NA"
    if isinstance(x, str):
        return x
    if isinstance(x, (bool, np.bool_)):
        return "TRUE" if x else "FALSE"
    if isinstance(x, (int, np.integer)):
        return str(int(x))
    x = float(x)
    if x != x:
        return "NA"
    if math.isinf(x):
        return "Inf" if x > 0 else "-Inf"
    if x == int(x) and abs(x) < 1e15:
        return str(int(x))
    s = f"{x:.{digits}g}"
    return s


def _fmt1(v):
    return " ".join(_strs(v)) if isinstance(v, V) else str(v)


# ======================================================================================== builtin implementations
def b_c(it, pos, named):
    parts = [(None, v) for v in pos] + list(named.items())
    parts = [(n, v) for n, v in parts if v is not None]
    if not parts:
        return None
    if any(isinstance(v, RList) for _, v in parts):
        items, names = [], []
        for n, v in parts:
            if isinstance(v, RList):
                items += v.items
                names += v.names or [""] * len(v.items)
            else:
                items += [V(v.flat()[i:i + 1]) for i in range(v.a.size)]
                names += v.names or [""] * v.a.size
        return RList(items, names if any(names) else None)
    arrs = [v.flat() for _, v in parts]
    kinds = {a.dtype.kind for a in arrs}
    if "O" in kinds:
        out = np.array([s if isinstance(s, str) or s is None else _fmt_elem(s, 15) for a in arrs for s in a],
                       dtype=object)
    elif "f" in kinds:
        out = np.concatenate([a.astype(np.float64) for a in arrs])
    elif "i" in kinds:
        out = np.concatenate([a.astype(np.int64) for a in arrs])
    else:
        out = np.concatenate(arrs)
    names = None
    if any(n or (v.names is not None and v.a.ndim == 1) for n, v in parts):
        names = []
        for n, v in parts:
            if v.names is not None and v.a.ndim == 1:
                names += [(f"{n}.{x}" if n else x) for x in v.names]
            elif n and v.a.size == 1:
                names.append(n)
            elif n:
                names += [f"{n}{i + 1}" for i in range(v.a.size)]
            else:
                names += [""] * v.a.size
    return V(out, names)


def b_list(it, pos, named):
    items = list(pos) + list(named.values())
    names = [""] * len(pos) + list(named.keys())
    return RList(items, names if named else None)


def b_length(it, pos, named):
    x = pos[0]
    return intv(0 if x is None else (len(x.exprs) if isinstance(x, Lang) else len(x)))


def b_dim(it, pos, named):
    x = pos[0]
    return intv(list(x.a.shape)) if isinstance(x, V) and x.a.ndim == 2 else None


def b_nrow(it, pos, named):
    x = pos[0]
    return intv(x.a.shape[0]) if isinstance(x, V) and x.a.ndim == 2 else None


def b_ncol(it, pos, named):
    x = pos[0]
    return intv(x.a.shape[1]) if isinstance(x, V) and x.a.ndim == 2 else None


def b_colon(it, pos, named):
    a, b = float(scalar(pos[0])), float(scalar(pos[1]))
    n = int(math.floor(abs(b - a) + 1e-10)) + 1
    step = 1 if b >= a else -1
    r = a + step * np.arange(n)
    return intv(r.astype(np.int64)) if a == int(a) else dbl(r)


def b_seq(it, pos, named):
    frm = _arg(pos, named, 0, "from")
    to = _arg(pos, named, 1, "to")
    by = _arg(pos, named, 2, "by")
    lo = named.get("length.out", named.get("length"))
    if frm is not None and to is None and by is None and lo is None:
        n = len(frm) if len(frm) != 1 else int(scalar(frm))
        return intv(np.arange(1, n + 1))
    a = float(scalar(frm)) if frm is not None else 1.0
    if lo is not None and to is None:
        n = int(scalar(lo))
        st = float(scalar(by)) if by is not None else 1.0
        return dbl(a + st * np.arange(n))
    b = float(scalar(to))
    if lo is not None:
        return dbl(np.linspace(a, b, int(scalar(lo))))
    st = float(scalar(by)) if by is not None else (1.0 if b >= a else -1.0)
    if (b - a) * st < 0:
        raise RError("wrong sign in 'by' argument")
    n = int(math.floor((b - a) / st + 1e-10)) + 1
    r = a + st * np.arange(n)
    allint = a == int(a) and st == int(st)
    return intv(r.astype(np.int64)) if allint else dbl(r)


def b_rep(it, pos, named):
    x = pos[0]
    times = _arg(pos, named, 1, "times")
    each = named.get("each")
    lo = named.get("length.out")
    if isinstance(x, RList):
        raise RError("rep() of lists is not supported")
    a = x.flat() if x is not None else np.zeros(0)
    if each is not None:
        a = np.repeat(a, int(scalar(each)))
    if times is not None:
        t = times.flat()
        a = np.tile(a, int(t[0])) if t.size == 1 else np.repeat(a, t.astype(np.int64))
    if lo is not None:
        a = np.resize(a, int(scalar(lo)))
    return V(a)


def b_matrix(it, pos, named):
    data = _arg(pos, named, 0, "data", dbl(np.nan))
    nrow = _arg(pos, named, 1, "nrow")
    ncol = _arg(pos, named, 2, "ncol")
    byrow = _flag(named.get("byrow"), False)
    d = data.flat()
    if nrow is None and ncol is None:
        nr, nc = d.size, 1
    elif nrow is None:
        nc = int(scalar(ncol))
        nr = -(-d.size // nc)
    elif ncol is None:
        nr = int(scalar(nrow))
        nc = -(-d.size // nr)
    else:
        nr, nc = int(scalar(nrow)), int(scalar(ncol))
    full = d if d.size == nr * nc else np.resize(d, nr * nc)
    m = full.reshape((nr, nc), order="C" if byrow else "F")
    return V(np.asfortranarray(m))


def b_t(it, pos, named):
    x = pos[0]
    if x.a.ndim == 1:
        return V(np.asfortranarray(x.a.reshape(1, -1)), None, [None, x.names] if x.names else None)
    dn = None if x.dimnames is None else [x.dimnames[1], x.dimnames[0]]
    return V(np.asfortranarray(x.a.T), None, dn)


def _as_mat(x, col=True):
    a = as_dbl(x) if x.a.dtype.kind != "f" else x.a
    if a.ndim == 2:
        return a
    return a.reshape(-1, 1) if col else a.reshape(1, -1)


def b_crossprod(it, pos, named):
    x = pos[0]
    y = pos[1] if len(pos) > 1 and pos[1] is not None else x
    a, b = _as_mat(x), _as_mat(y)
    if a.shape[0] != b.shape[0]:
        raise RError("non-conformable arguments")
    dn = [x.dimnames[1] if x.a.ndim == 2 and x.dimnames else None, y.dimnames[1] if y.a.ndim == 2 and y.dimnames else None]
    return V(np.asfortranarray(a.T @ b), None, dn if any(d is not None for d in dn) else None)


def b_tcrossprod(it, pos, named):
    x = pos[0]
    y = pos[1] if len(pos) > 1 and pos[1] is not None else x
    a, b = _as_mat(x), _as_mat(y)
    if a.shape[1] != b.shape[1]:
        raise RError("non-conformable arguments")
    return V(np.asfortranarray(a @ b.T))


def b_matmul(it, pos, named):
    x, y = pos
    a = as_dbl(x)
    b = as_dbl(y)
    if a.ndim == 1 and b.ndim == 1:
        if a.size == b.size:
            return V(np.asfortranarray(np.array([[a @ b]])))
        if a.size == 1:
            return V(np.asfortranarray(a[0] * b.reshape(1, -1)))
        raise RError("non-conformable arguments")
    if a.ndim == 1:
        a = a.reshape(1, -1) if a.size == b.shape[0] else a.reshape(-1, 1)
    if b.ndim == 1:
        b = b.reshape(-1, 1) if b.size == a.shape[1] else b.reshape(1, -1)
    if a.shape[1] != b.shape[0]:
        raise RError("non-conformable arguments")
    rn = x.dimnames[0] if x.a.ndim == 2 and x.dimnames else None
    cn = y.dimnames[1] if y.a.ndim == 2 and y.dimnames else None
    return V(np.asfortranarray(a @ b), None, [rn, cn] if rn is not None or cn is not None else None)


def _narm(named):
    return _flag(named.get("na.rm"), False)


def b_sum(it, pos, named):
    tot = LD(0)
    allint = True
    for v in pos:
        if v is None:
            continue
        a = as_num(v).reshape(-1)
        if a.dtype.kind == "f":
            allint = False
            if _narm(named):
                a = a[~np.isnan(a)]
        tot += np.sum(a.astype(LD))
    return intv(int(tot)) if allint else dbl(np.float64(tot))


def b_prod(it, pos, named):
    return dbl(np.prod(np.concatenate([as_dbl(v).reshape(-1) for v in pos]).astype(LD)).astype(np.float64))


def _minmax(fn):
    def f(it, pos, named):
        arrs = [as_num(v).reshape(-1) for v in pos if v is not None]
        a = np.concatenate(arrs) if arrs else np.zeros(0)
        if _narm(named) and a.dtype.kind == "f":
            a = a[~np.isnan(a)]
        if a.size == 0:
            it.warn("no non-missing arguments to max/min; returning -Inf/Inf")
            return dbl(-np.inf if fn is np.max else np.inf)
        r = fn(a)  # NaN propagates like NA
        return intv(r) if a.dtype.kind == "i" else dbl(r)
    return f


def b_mean(it, pos, named):
    a = as_dbl(pos[0]).reshape(-1)
    if _narm(named):
        a = a[~np.isnan(a)]
    if a.size == 0:
        return dbl(np.nan)
    m = np.sum(a.astype(LD)) / a.size
    m = m + np.sum(a.astype(LD) - m) / a.size  # R's second pass (summary.c)
    return dbl(np.float64(m))


def b_var(it, pos, named):
    a = as_dbl(pos[0]).reshape(-1)
    if _narm(named):
        a = a[~np.isnan(a)]
    n = a.size
    if n < 2:
        return dbl(np.nan)
    al = a.astype(LD)
    m = np.sum(al) / n
    m = m + np.sum(al - m) / n
    return dbl(np.float64(np.sum((al - m) ** 2) / (n - 1)))


def b_median(it, pos, named):
    a = as_dbl(pos[0]).reshape(-1)
    if _narm(named):
        a = a[~np.isnan(a)]
    elif np.isnan(a).any():
        return dbl(np.nan)
    if a.size == 0:
        return dbl(np.nan)
    s = np.sort(a)
    h = (a.size + 1) // 2
    return dbl(s[h - 1] if a.size % 2 == 1 else (s[h - 1] + s[h]) / 2)  # mean(sort(x, partial = half + 0:1)[half + 0:1])


def _colrow(axis, mean=False):
    def f(it, pos, named):
        x = pos[0]
        if not isinstance(x, V) or x.a.ndim != 2:
            raise RError("'x' must be an array of at least two dimensions")
        a = as_num(x).astype(LD)
        if _narm(named):
            nan = np.isnan(a)
            cnt = (~nan).sum(axis=axis)
            a = np.where(nan, 0, a)
        else:
            cnt = a.shape[axis]
        s = np.sum(a, axis=axis)
        if mean:
            s = s / cnt
        dn = x.dimnames[1 - axis] if x.dimnames else None
        return V(s.astype(np.float64), dn)
    return f


def b_cumsum(it, pos, named):
    x = pos[0]
    a = as_num(x).reshape(-1, order="F")
    if a.dtype.kind == "i":
        return V(np.cumsum(a), x.names)
    return V(np.cumsum(a.astype(LD)).astype(np.float64), x.names)


def b_any(it, pos, named):
    r = False
    for v in pos:
        if v is None:
            continue
        a = v.a
        r = r or bool(np.any(a.astype(bool) if a.dtype.kind != "f" else (a != 0) & ~np.isnan(a)))
    return lgl(r)


def b_all(it, pos, named):
    r = True
    for v in pos:
        if v is None:
            continue
        r = r and bool(np.all(v.a.astype(bool)))
    return lgl(r)


def b_which(it, pos, named):
    x = pos[0]
    k = np.flatnonzero(x.flat().astype(bool))
    return V(k + 1, None if x.names is None or x.a.ndim == 2 else [x.names[i] for i in k])


def b_ifelse(it, pos, named):
    test = _arg(pos, named, 0, "test")
    yes = _arg(pos, named, 1, "yes")
    no = _arg(pos, named, 2, "no")
    t = test.a.astype(bool)
    n = t.size
    tf = t.reshape(-1, order="F")
    ya, na_ = yes.flat(), no.flat()
    kind = "O" if "O" in (ya.dtype.kind, na_.dtype.kind) else ("f" if "f" in (ya.dtype.kind, na_.dtype.kind) else
                                                                 ya.dtype.kind)
    out = np.empty(n, dtype={"O": object, "f": np.float64, "i": np.int64, "b": bool}[kind])
    if tf.any():
        out[tf] = np.resize(ya, n)[tf]
    if (~tf).any():
        out[~tf] = np.resize(na_, n)[~tf]
    if t.ndim == 2:
        return V(out.reshape(t.shape, order="F"), None, test.dimnames)
    return V(out, test.names)


def b_is_na(it, pos, named):
    x = pos[0]
    if isinstance(x, RList):
        return lgl([False] * len(x.items))
    if x is None:
        return lgl(np.zeros(0, bool))
    a = x.a
    if a.dtype.kind == "f":
        r = np.isnan(a)
    elif a.dtype.kind == "O":
        r = np.array([s is None for s in a.reshape(-1)]).reshape(a.shape)
    else:
        r = np.zeros(a.shape, dtype=bool)
    return V(r, x.names, x.dimnames)


def b_sweep(it, pos, named):
    x = _arg(pos, named, 0, "x")
    margin = int(scalar(_arg(pos, named, 1, "MARGIN")))
    stats = _arg(pos, named, 2, "STATS")
    fun = _arg(pos, named, 3, "FUN", it.lookup_fn("-", it.globalenv))
    if isinstance(fun, V):
        fun = it.lookup_fn(fun.a[0], it.globalenv)
    s = stats.flat()
    ext = x.a.shape[margin - 1]
    if s.size != ext:
        if s.size == 0 or ext % s.size:
            it.warn("STATS does not recycle exactly across MARGIN")  # check.margin = TRUE
        s = np.resize(s, ext)
    # aperm(array(STATS, dims[perm]), order(perm)): STATS laid along MARGIN, constant along the other
    full = np.repeat(s.reshape(-1, 1), x.a.shape[1], axis=1) if margin == 1 else np.repeat(s.reshape(1, -1),
                                                                                          x.a.shape[0], axis=0)
    return it.call_value(fun, [x, V(np.asfortranarray(full))])


def b_apply(it, pos, named):
    x = _arg(pos, named, 0, "X")
    margin = int(scalar(_arg(pos, named, 1, "MARGIN")))
    fun = _arg(pos, named, 2, "FUN")
    extra = list(pos[3:])
    kw = {k: v for k, v in named.items() if k not in ("X", "MARGIN", "FUN")}
    res = []
    n = x.a.shape[margin - 1]
    dn = x.dimnames or [None, None]
    for i in range(n):
        sl = x.a[i, :] if margin == 1 else x.a[:, i]
        res.append(it.call_value(fun, [V(np.ascontiguousarray(sl), dn[2 - margin])] + extra, kw))
    return _simplify(res, dn[margin - 1])


def _simplify(res, names=None):
    if not res:
        return RList([])
    if all(isinstance(r, V) and r.a.size == 1 for r in res):
        out = b_c(None, [V(r.flat()) for r in res], {})
        out.names = names
        return out
    if all(isinstance(r, V) for r in res) and len({r.a.size for r in res}) == 1 and res[0].a.size > 0:
        cols = [r.flat() for r in res]
        m = np.stack(cols, axis=1)
        return V(np.asfortranarray(m), None, [res[0].names if res[0].a.ndim == 1 else None, names]
                 if (names is not None or res[0].names is not None) else None)
    return RList(res, names)


def _iter_items(x):
    if x is None:
        return [], None
    if isinstance(x, RList):
        return x.items, x.names
    return [V(x.flat()[i:i + 1]) for i in range(x.a.size)], x.names


def b_lapply(it, pos, named):
    x = _arg(pos, named, 0, "X")
    fun = _arg(pos, named, 1, "FUN")
    extra = list(pos[2:])
    kw = {k: v for k, v in named.items() if k not in ("X", "FUN")}
    items, names = _iter_items(x)
    return RList([it.call_value(fun, [v] + extra, kw) for v in items], names)


def b_sapply(it, pos, named):
    x = _arg(pos, named, 0, "X")
    fun = _arg(pos, named, 1, "FUN")
    extra = list(pos[2:])
    kw = {k: v for k, v in named.items() if k not in ("X", "FUN", "simplify", "USE.NAMES")}
    items, names = _iter_items(x)
    if names is None and isinstance(x, V) and x.a.dtype.kind == "O":
        names = list(x.flat())
    return _simplify([it.call_value(fun, [v] + extra, kw) for v in items], names)


def b_unlist(it, pos, named):
    x = pos[0]
    if not isinstance(x, RList):
        return x
    flat = []

    def rec(v, prefix):
        if isinstance(v, RList):
            for i, s in enumerate(v.items):
                n = v.names[i] if v.names else ""
                rec(s, n or prefix)
        elif v is not None:
            flat.append((prefix, v))
    rec(x, "")
    if not flat:
        return None
    use_names = _flag(named.get("use.names"), True)
    out = b_c(it, [v for _, v in flat], {})
    if use_names and any(n for n, _ in flat) and out.names is None:
        names = []
        for n, v in flat:
            names += [n if v.a.size == 1 else f"{n}{i + 1}" for i in range(v.a.size)]
        out.names = names
    return out


def b_names(it, pos, named):
    x = pos[0]
    if isinstance(x, RList):
        return chr_(x.names) if x.names is not None else None
    if isinstance(x, V):
        if x.a.ndim == 2:
            return chr_(x.dimnames[1]) if x.dimnames and x.dimnames[1] is not None else None
        return chr_(x.names) if x.names is not None else None
    return None


def b_names_assign(it, pos, named):
    x, val = pos[0], named["value"]
    nm = None if val is None else [s for s in _strs(val)]
    if isinstance(x, RList):
        return RList(x.items, nm, x.attrs)
    if nm is not None and len(nm) != x.a.size:
        nm = (nm + [None] * x.a.size)[:x.a.size]
    return V(x.a, nm, x.dimnames, x.attrs)


def _dimnames_get(k):
    def f(it, pos, named):
        x = pos[0]
        if isinstance(x, V) and x.a.ndim == 2 and x.dimnames and x.dimnames[k] is not None:
            return chr_(x.dimnames[k])
        return None
    return f


def _dimnames_set(k):
    def f(it, pos, named):
        x, val = pos[0], named["value"]
        if not isinstance(x, V) or x.a.ndim != 2:
            raise RError("attempt to set 'rownames'/'colnames' on an object with no dimensions")
        dn = list(x.dimnames) if x.dimnames else [None, None]
        if val is None:
            dn[k] = None
        else:
            nm = _strs(val) if val.a.dtype.kind != "O" else list(val.flat())
            if len(nm) != x.a.shape[k]:
                raise RError(f"length of 'dimnames' [{k + 1}] not equal to array extent")
            dn[k] = nm
        return V(x.a, None, dn if any(d is not None for d in dn) else None, x.attrs)
    return f


def b_dimnames(it, pos, named):
    x = pos[0]
    if isinstance(x, V) and x.a.ndim == 2 and x.dimnames is not None:
        return RList([None if d is None else chr_(d) for d in x.dimnames])
    return None


def b_dimnames_assign(it, pos, named):
    x, val = pos[0], named["value"]
    if not isinstance(x, V) or x.a.ndim != 2:
        raise RError("'dimnames' applied to non-array")
    if val is None:
        return V(x.a, None, None, x.attrs)
    dn = [None if d is None else (list(d.flat()) if d.a.dtype.kind == "O" else _strs(d)) for d in val.items]
    for k in (0, 1):
        if dn[k] is not None and len(dn[k]) != x.a.shape[k]:
            raise RError(f"length of 'dimnames' [{k + 1}] not equal to array extent")
    return V(x.a, None, dn if any(d is not None for d in dn) else None, x.attrs)


def b_setNames(it, pos, named):
    x = _arg(pos, named, 0, "object")
    nm = _arg(pos, named, 1, "nm")
    return b_names_assign(it, [x], {"value": nm})


def b_class(it, pos, named):
    x = pos[0]
    attrs = getattr(x, "attrs", None)
    if attrs and "class" in attrs:
        return attrs["class"]
    if isinstance(x, RList):
        return chr_("list")
    if isinstance(x, (Closure, Builtin)):
        return chr_("function")
    if x is None:
        return chr_("NULL")
    if x.a.ndim == 2:
        return chr_(["matrix", "array"])
    return chr_({"double": "numeric"}.get(x.kind, x.kind))


def b_class_assign(it, pos, named):
    x, val = pos[0], named["value"]
    attrs = dict(getattr(x, "attrs", None) or {})
    attrs["class"] = val
    if isinstance(x, RList):
        return RList(x.items, x.names, attrs)
    return V(x.a, x.names, x.dimnames, attrs)


def b_inherits(it, pos, named):
    x, what = pos[0], pos[1]
    cl = set(b_class(it, [x], {}).flat())
    return lgl(bool(cl & set(what.flat())))


def b_attr(it, pos, named):
    x, which = pos[0], pos[1].a[0]
    attrs = getattr(x, "attrs", None) or {}
    return attrs.get(which)


def b_attr_assign(it, pos, named):
    x, which, val = pos[0], pos[1].a[0], named["value"]
    attrs = dict(getattr(x, "attrs", None) or {})
    attrs[which] = val
    if isinstance(x, RList):
        return RList(x.items, x.names, attrs)
    return V(x.a, x.names, x.dimnames, attrs)


def b_as_vector(it, pos, named):
    x = pos[0]
    if x is None or isinstance(x, RList):
        return x
    return V(x.flat().copy() if x.a.ndim == 2 else x.a)


def b_as_numeric(it, pos, named):
    x = pos[0]
    if x is None:
        return dbl(np.zeros(0))
    return V(as_dbl(x).reshape(-1, order="F").astype(np.float64))


def b_as_integer(it, pos, named):
    x = pos[0]
    return V(np.trunc(as_dbl(x).reshape(-1, order="F")).astype(np.int64), x.names if isinstance(x, V) else None)


def b_as_character(it, pos, named):
    x = pos[0]
    if isinstance(x, Lang):
        return chr_([P.deparse(e) for e in x.exprs])
    if x is None:
        return V(np.empty(0, dtype=object))
    return chr_([s if isinstance(s, str) else _fmt_elem(s, 15) for s in (x.flat() if isinstance(x, V) else _strs(x))])


def b_as_logical(it, pos, named):
    return V(as_num(pos[0]).astype(bool))


def b_as_matrix(it, pos, named):
    x = pos[0]
    if x.a.ndim == 2:
        return x
    return V(np.asfortranarray(x.a.reshape(-1, 1)), None, [x.names, None] if x.names else None)


def b_all_equal(it, pos, named):
    target, current = pos[0], pos[1]
    tol = float(scalar(named["tolerance"])) if "tolerance" in named else 1.5e-8
    if not (isinstance(target, V) and isinstance(current, V)):
        raise RError("all.equal(): only numeric vectors are supported")
    t, c = as_dbl(target).reshape(-1), as_dbl(current).reshape(-1)
    if t.size != c.size:
        return chr_(f"Lengths ({t.size}, {c.size}) differ")
    out = np.isnan(t) | np.isnan(c)
    if (np.isnan(t) != np.isnan(c)).any():
        return chr_("'is.NA' value mismatch")
    t, c = t[~out], c[~out]
    n = t.size
    if n == 0:
        return lgl(True)
    xy = np.sum(np.abs(t - c)) / n
    what = "absolute"
    xn = np.sum(np.abs(t)) / n
    if np.isfinite(xn) and xn > tol:
        xy = xy / xn
        what = "relative"
    if np.isnan(xy) or xy > tol:
        return chr_(f"Mean {what} difference: {xy:.7g}")
    return lgl(True)


def b_isTRUE(it, pos, named):
    x = pos[0]
    return lgl(isinstance(x, V) and x.a.dtype.kind == "b" and x.a.size == 1 and bool(x.a.reshape(-1)[0]))


def b_identical(it, pos, named):
    x, y = pos[0], pos[1]
    return lgl(_identical(x, y))


def _identical(x, y):
    if x is None or y is None:
        return x is None and y is None
    if isinstance(x, V) and isinstance(y, V):
        return x.a.shape == y.a.shape and x.a.dtype.kind == y.a.dtype.kind and bool(
            np.all((x.a == y.a) | ((x.a != x.a) & (y.a != y.a))) if x.a.dtype.kind == "f" else np.all(x.a == y.a))
    if isinstance(x, RList) and isinstance(y, RList):
        return len(x.items) == len(y.items) and all(_identical(a, b) for a, b in zip(x.items, y.items))
    return x is y


def b_order(it, pos, named):
    x = pos[0]
    dec = _flag(named.get("decreasing"), False)
    a = as_num(x).reshape(-1, order="F")
    if a.dtype.kind == "f":
        nan = np.isnan(a)
        key = np.where(nan, np.inf, -a if dec else a)  # NA last (na.last = TRUE), ties keep their original order
        idx = np.lexsort((np.arange(a.size), nan, key)) if nan.any() else np.argsort(key, kind="stable")
    else:
        idx = np.argsort(-a if dec else a, kind="stable")
    return intv(idx + 1)


def b_sort(it, pos, named):
    x = pos[0]
    dec = _flag(named.get("decreasing"), False)
    a = x.flat()
    if a.dtype.kind == "f":
        a = a[~np.isnan(a)]
    idx = np.argsort(-as_num(V(a)) if dec else a, kind="stable")
    return V(a[idx], None if x.names is None else [x.names[i] for i in idx])


def b_rev(it, pos, named):
    x = pos[0]
    return V(x.flat()[::-1].copy(), None if x.names is None else x.names[::-1])


def b_unique(it, pos, named):
    a = pos[0].flat()
    seen, out = set(), []
    for v in a:
        key = v if not (isinstance(v, float) and v != v) else "NaN"
        if key not in seen:
            seen.add(key)
            out.append(v)
    return V(np.array(out, dtype=a.dtype))


def _col_key(col):
    c = np.array(col, dtype=np.float64, copy=True)
    c[c == 0] = 0.0  # -0 == 0
    c[np.isnan(c)] = np.nan  # one NaN payload
    return c.tobytes()


def b_duplicated(it, pos, named):
    x = pos[0]
    margin = int(scalar(named["MARGIN"])) if "MARGIN" in named else 1
    from_last = _flag(named.get("fromLast"), False)
    if x.a.ndim == 2:
        n = x.a.shape[margin - 1]
        keys = [_col_key(x.a[:, j] if margin == 2 else x.a[j, :]) for j in range(n)]
    else:
        keys = [v if not (isinstance(v, float) and v != v) else "NaN" for v in x.flat()]
    out = np.zeros(len(keys), dtype=bool)
    seen = set()
    rng = range(len(keys) - 1, -1, -1) if from_last else range(len(keys))
    for j in rng:
        if keys[j] in seen:
            out[j] = True
        else:
            seen.add(keys[j])
    return lgl(out)


def b_data_frame(it, pos, named):
    x = pos[0]
    if x.a.ndim != 2:
        raise RError("data.frame(): only matrices are supported")
    cn = x.dimnames[1] if x.dimnames and x.dimnames[1] is not None else [f"X{j + 1}" for j in range(x.a.shape[1])]
    return RList([V(np.ascontiguousarray(x.a[:, j])) for j in range(x.a.shape[1])], cn, {"class": chr_("data.frame")})


def b_match(it, pos, named):
    x, table = pos[0], pos[1]
    if isinstance(x, RList) or isinstance(table, RList):
        tk = [_col_key(v.a) for v in table.items]
        out = []
        for v in x.items:
            k = _col_key(v.a)
            out.append(tk.index(k) + 1 if k in tk else np.nan)
        return V(np.array(out, dtype=np.float64)) if any(o != o for o in out) else intv(out)
    lut = {}
    for i, v in enumerate(table.flat() if table is not None else []):
        lut.setdefault(v, i + 1)
    out = [lut.get(v, np.nan) for v in (x.flat() if x is not None else [])]
    return V(np.array(out, dtype=np.float64)) if any(o != o for o in out) else intv(out)


def b_in(it, pos, named):
    x, table = pos
    if x is None:
        return lgl(np.zeros(0, bool))
    t = set(table.flat().tolist()) if table is not None else set()
    return V(np.array([v in t for v in x.flat().tolist()], dtype=bool))


def b_scale(it, pos, named):
    x = _arg(pos, named, 0, "x")
    center = _arg(pos, named, 1, "center", lgl(True))
    scl = _arg(pos, named, 2, "scale", lgl(True))
    a = as_dbl(x).astype(np.float64).copy(order="F")
    if a.ndim != 2:
        a = a.reshape(-1, 1)
    attrs = {}
    if center.a.dtype.kind == "b":
        if truthy(center):
            nan = np.isnan(a)
            cm = (np.sum(np.where(nan, 0, a).astype(LD), axis=0) / (~nan).sum(axis=0)).astype(np.float64)  # colMeans(na.rm)
            a = a - cm[None, :]
            attrs["scaled:center"] = V(cm)
    else:
        a = a - as_dbl(center).reshape(1, -1)
    if scl.a.dtype.kind == "b":
        if truthy(scl):
            nan = np.isnan(a)
            ss = np.sum(np.where(nan, 0, a * a).astype(LD), axis=0).astype(np.float64)
            cnt = (~nan).sum(axis=0)
            sd = np.sqrt(ss / np.maximum(1, cnt - 1))  # sqrt(sum(v^2) / max(1, length(v) - 1L))
            with np.errstate(all="ignore"):
                a = a / sd[None, :]
            attrs["scaled:scale"] = V(sd)
    else:
        a = a / as_dbl(scl).reshape(1, -1)
    return V(np.asfortranarray(a), None, x.dimnames, attrs)


def b_pnorm(it, pos, named):
    q = _arg(pos, named, 0, "q")
    mean = as_dbl(_arg(pos, named, 1, "mean", dbl(0.0)))
    sd = as_dbl(_arg(pos, named, 2, "sd", dbl(1.0)))
    lower = _flag(named.get("lower.tail", pos[3] if len(pos) > 3 else None), True)
    logp = _flag(named.get("log.p", pos[4] if len(pos) > 4 else None), False)
    z = as_dbl(q)
    if not (mean.size == 1 and mean[0] == 0 and sd.size == 1 and sd[0] == 1):
        z = (z - (mean if mean.size == 1 else mean.reshape(z.shape))) / (sd if sd.size == 1 else sd.reshape(z.shape))
    if not lower:
        z = -z
    r = sp.log_ndtr(z) if logp else sp.ndtr(z)
    return V(np.asarray(r, dtype=np.float64), q.names, q.dimnames)


def b_qnorm(it, pos, named):
    p = _arg(pos, named, 0, "p")
    r = sp.ndtri(as_dbl(p))
    return V(np.asarray(r), p.names, p.dimnames)


def b_gamma_inc(it, pos, named):
    a, x = as_dbl(pos[0]), as_dbl(pos[1])
    if (a <= 0).any():
        raise RError("gsl::gamma_inc(a, x) with a <= 0 is not supported by this stand-in")
    with np.errstate(all="ignore"):
        r = sp.gamma(a) * sp.gammaincc(a, x)
    src = pos[1] if x.size >= a.size else pos[0]
    return V(np.atleast_1d(np.asarray(r, dtype=np.float64)), src.names, src.dimnames)


def b_expint_E1(it, pos, named):
    x = pos[0]
    return V(np.atleast_1d(sp.exp1(as_dbl(x))), x.names, x.dimnames)


def b_hyperg_1F1(it, pos, named):
    a, b, x = (as_dbl(v) for v in pos[:3])
    return V(np.atleast_1d(sp.hyp1f1(a, b, x)))


def b_OwensT(it, pos, named):
    h, a = as_dbl(pos[0]), as_dbl(pos[1])
    return V(np.atleast_1d(sp.owens_t(h, a)))


def b_paste(sep_default):
    def f(it, pos, named):
        sep = named["sep"].a[0] if "sep" in named else sep_default
        collapse = named["collapse"].a[0] if named.get("collapse") is not None else None
        parts = [_strs(v) for v in pos]
        parts = [p for p in parts if len(p) > 0]
        if not parts:
            return chr_([""]) if collapse is not None else V(np.empty(0, dtype=object))
        n = max(len(p) for p in parts)
        out = [sep.join(p[i % len(p)] for p in parts) for i in range(n)]
        if collapse is not None:
            out = [collapse.join(out)]
        return chr_(out)
    return f


def b_format(it, pos, named):
    x = pos[0]
    digits = int(scalar(named["digits"])) if "digits" in named else 7
    return chr_([_fmt_elem(v, digits) for v in (x.flat() if isinstance(x, V) else [])])


def b_cat(it, pos, named):
    sep = named["sep"].a[0] if "sep" in named else " "
    it.out.append(sep.join(s for v in pos for s in _strs(v)))
    return None


def b_print(it, pos, named):
    it.out.append(repr(pos[0]))
    return pos[0]


def b_stop(it, pos, named):
    raise RError("".join(s for v in pos for s in _strs(v)))


def b_warning(it, pos, named):
    it.warn("".join(s for v in pos for s in _strs(v)))
    return None


def b_stopifnot(it, pos, named):
    for v in pos:
        if not (isinstance(v, V) and v.a.size > 0 and bool(np.all(v.a.astype(bool)))):
            raise RError("stopifnot(): condition is not all TRUE")
    return None


def b_is(kind):
    def f(it, pos, named):
        x = pos[0]
        if kind == "null":
            return lgl(x is None)
        if kind == "list":
            return lgl(isinstance(x, RList))
        if kind == "function":
            return lgl(isinstance(x, (Closure, Builtin)))
        if not isinstance(x, V):
            return lgl(kind == "vector" and isinstance(x, RList))
        k = x.a.dtype.kind
        if kind == "vector":  # no attributes other than names
            return lgl(x.a.ndim == 1 and not x.attrs)
        if kind == "matrix":
            return lgl(x.a.ndim == 2)
        return lgl({"numeric": k in "fi", "double": k == "f", "integer": k == "i", "logical": k == "b",
                    "character": k == "O"}[kind])
    return f


def b_is_elementwise(fn):
    def f(it, pos, named):
        x = pos[0]
        if x.a.dtype.kind != "f":
            r = np.full(x.a.shape, fn is np.isfinite and x.a.dtype.kind in "ib")
        else:
            r = fn(x.a)
        return V(r, x.names, x.dimnames)
    return f


def b_round(it, pos, named):
    x = pos[0]
    d = int(scalar(_arg(pos, named, 1, "digits", intv(0))))
    return V(np.round(as_dbl(x), d), x.names, x.dimnames)


def b_diff(it, pos, named):
    return V(np.diff(as_num(pos[0]).reshape(-1)))


def b_cbind(axis):
    def f(it, pos, named):
        cols, names = [], []
        for n, v in [(None, v) for v in pos] + list(named.items()):
            if v is None:
                continue
            a = v.a
            if a.ndim == 1:
                a = a.reshape(-1, 1) if axis == 1 else a.reshape(1, -1)
                names.append([n or ""])
            else:
                dn = v.dimnames[axis] if v.dimnames and v.dimnames[axis] is not None else [""] * a.shape[axis]
                names.append(dn)
            cols.append(a)
        m = np.concatenate(cols, axis=axis)
        flat = [x for n in names for x in n]
        dn = [None, None]
        if any(flat):
            dn[axis] = flat
        return V(np.asfortranarray(m), None, dn if any(d is not None for d in dn) else None)
    return f


def b_uniroot(it, pos, named):
    """stats::uniroot -> R_zeroin2 (Brent's zeroin, the Netlib C translation R ships), default
    tol = .Machine$double.eps^0.25, maxiter = 1000."""
    f = _arg(pos, named, 0, "f")
    interval = _arg(pos, named, 1, "interval")
    lower = float(scalar(named["lower"])) if "lower" in named else float(interval.flat()[0])
    upper = float(scalar(named["upper"])) if "upper" in named else float(interval.flat()[1])
    tol = float(scalar(named["tol"])) if "tol" in named else np.finfo(np.float64).eps ** 0.25
    maxit = int(scalar(named["maxiter"])) if "maxiter" in named else 1000

    def fn(x):
        return float(scalar(it.call_value(f, [dbl(x)])))
    a, b = lower, upper
    fa, fb = fn(a), fn(b)
    if not (np.isfinite(fa) and np.isfinite(fb)):
        raise RError("f.lower / f.upper = f(lower / upper) is NA or infinite")
    if fa * fb > 0:
        raise RError("f() values at end points not of opposite sign")
    EPS = np.finfo(np.float64).eps
    c, fc = a, fa
    iters = maxit + 1
    est = 0.0
    if fa == 0.0:
        b, fb, iters_used = a, fa, 0
    elif fb == 0.0:
        iters_used = 0
    else:
        iters_used = -1
        while iters > 0:
            iters -= 1
            prev_step = b - a
            if abs(fc) < abs(fb):
                a, b, c = b, c, b
                fa, fb, fc = fb, fc, fb
            tol_act = 2 * EPS * abs(b) + tol / 2
            new_step = (c - b) / 2
            if abs(new_step) <= tol_act or fb == 0.0:
                iters_used = maxit + 1 - iters - 1
                est = abs(c - b)
                break
            if abs(prev_step) >= tol_act and abs(fa) > abs(fb):
                cb = c - b
                if a == c:
                    t1 = fb / fa
                    p = cb * t1
                    q = 1.0 - t1
                else:
                    q = fa / fc
                    t1 = fb / fc
                    t2 = fb / fa
                    p = t2 * (cb * q * (q - t1) - (b - a) * (t1 - 1.0))
                    q = (q - 1.0) * (t1 - 1.0) * (t2 - 1.0)
                if p > 0:
                    q = -q
                else:
                    p = -p
                if p < (0.75 * cb * q - abs(tol_act * q) / 2) and p < abs(prev_step * q / 2):
                    new_step = p / q
            if abs(new_step) < tol_act:
                new_step = tol_act if new_step > 0 else -tol_act
            a, fa = b, fb
            b += new_step
            fb = fn(b)
            if (fb > 0 and fc > 0) or (fb < 0 and fc < 0):
                c, fc = a, fa
        if iters_used < 0:
            it.warn("_NOT_ converged in maxiter iterations")
            iters_used = maxit
    return RList([dbl(b), dbl(fn(b)), intv(iters_used), dbl(est)], ["root", "f.root", "iter", "estim.prec"])


# ---------------------------------------------------------------------------------------- special forms
def sp_with(it, env, args):
    data = it.eval(args[0][1], env)
    wenv = Env(env)
    if isinstance(data, RList):
        for n, v in zip(data.names or [], data.items):
            if n:
                wenv.vars[n] = v
    return it.eval(args[1][1], wenv)


def sp_quote(it, env, args):
    return Lang([args[0][1]])


def _arg_exprs(it, env):
    """formal -> argument expression of the closure call that created env."""
    e = env
    while e is not None and e.fn is None:
        e = e.parent
    if e is None or e.call is None:
        return {}
    formals = [p[0] for p in e.fn.params]
    out, rest = {}, []
    for name, ex in e.call[2]:
        if name is not None and name in formals:
            out[name] = ex
        elif name is None:
            rest.append(ex)
    free = [f for f in formals if f not in out and f != "..."]
    for f, ex in zip(free, rest):
        out[f] = ex
    return out


def sp_substitute(it, env, args):
    ex = args[0][1]
    if ex[0] == "id":
        return Lang([_arg_exprs(it, env).get(ex[1], ex)])
    return Lang([ex])


def sp_missing(it, env, args):
    name = args[0][1][1]
    e = env
    while e is not None and ".supplied" not in e.vars:
        e = e.parent
    return lgl(e is None or name not in e.vars[".supplied"])


def sp_match_call(it, env, args):
    e = env
    while e is not None and e.fn is None:
        e = e.parent
    return it.match_call(e)


def sp_rm(it, env, args):
    for name, ex in args:
        if name is None and ex[0] == "id":
            env.vars.pop(ex[1], None)
    return None


def sp_tryCatch(it, env, args):
    handlers = {n: ex for n, ex in args if n is not None}
    body = [ex for n, ex in args if n is None]
    try:
        val = None
        for ex in body:
            val = it.eval(ex, env)
        return val
    except RError as err:
        if "error" in handlers:
            h = it.eval(handlers["error"], env)
            return it.call_value(h, [RList([chr_(str(err))], ["message"], {"class": chr_(["simpleError", "error"])})])
        raise
    finally:
        if "finally" in handlers:
            it.eval(handlers["finally"], env)


def sp_return(it, env, args):
    from .interp import ReturnEx
    raise ReturnEx(it.eval(args[0][1], env) if args else None)


def sp_function_noop(it, env, args):
    return None


def sp_suppress(it, env, args):
    n = len(it.warnings)
    v = it.eval(args[0][1], env)
    del it.warnings[n:]
    return v


def sp_save(it, env, args):
    """save(obj1, obj2, ..., file = path): the named objects, as an .npz archive standing in for R's .RData format
    (lists are flattened to `name$field` entries)."""
    path = None
    objs = {}
    for name, ex in args:
        if name == "file":
            path = it.eval(ex, env).a[0]
        elif name is None and ex[0] == "id":
            objs[ex[1]] = it.eval(ex, env)
    if path is None:
        raise RError("save(): 'file' must be specified")
    flat = {}

    def put(key, v):
        if isinstance(v, RList):
            for i, item in enumerate(v.items):
                put(f"{key}${v.names[i] if v.names and v.names[i] else i + 1}", item)
        elif isinstance(v, V):
            flat[key] = v.a if v.a.dtype.kind != "O" else np.array([str(x) for x in v.a.reshape(-1)])
        elif v is None:
            flat[key] = np.zeros(0)
    for k, v in objs.items():
        put(k, v)
    with open(path, "wb") as f:
        np.savez(f, **flat)
    return None


def b_list_files(it, pos, named):
    import re
    path = _arg(pos, named, 0, "path", chr_(".")).a[0]
    pattern = _arg(pos, named, 1, "pattern")
    names = sorted(os.listdir(path)) if os.path.isdir(path) else []
    if pattern is not None:
        rx = re.compile(pattern.a[0])
        names = [n for n in names if rx.search(n)]
    return chr_(names) if names else V(np.empty(0, dtype=object))


def b_file_remove(it, pos, named):
    out = []
    for f in _strs(pos[0]):
        try:
            os.remove(f)
            out.append(True)
        except OSError:
            out.append(False)
    return lgl(out)


def sp_switch(it, env, args):
    sel = it.eval(args[0][1], env)
    alts = args[1:]
    if sel.a.dtype.kind == "O":
        key = sel.a[0]
        for i, (n, ex) in enumerate(alts):
            if n == key:
                j = i
                while alts[j][1] is None:
                    j += 1
                return it.eval(alts[j][1], env)
        for n, ex in alts:
            if n is None and ex is not None:
                return it.eval(ex, env)
        return None
    k = int(scalar(sel)) - 1
    return it.eval(alts[k][1], env) if 0 <= k < len(alts) else None


def b_deparse(it, pos, named):
    x = pos[0]
    if isinstance(x, Lang):
        return chr_([P.deparse(x.exprs[0])])
    return chr_([_fmt1(x)])


def b_invisible(it, pos, named):
    return pos[0] if pos else None


def b_do_call(it, pos, named):
    what, args = pos[0], pos[1]
    fn = it.lookup_fn(what.a[0], it.globalenv) if isinstance(what, V) else what
    p = [v for n, v in zip(args.names or [""] * len(args.items), args.items) if not n]
    k = {n: v for n, v in zip(args.names or [], args.items) if n}
    return it.call_value(fn, p, k)


def b_file_path(it, pos, named):
    parts = [_strs(v) for v in pos]
    return chr_([os.path.join(*[p[0] for p in parts])])


def b_nchar(it, pos, named):
    return intv([len(s) for s in _strs(pos[0])])


def b_numeric(it, pos, named):
    return dbl(np.zeros(int(scalar(pos[0])) if pos else 0))


def b_diag(it, pos, named):
    x = pos[0]
    if x.a.ndim == 2:
        return V(np.diag(x.a).copy())
    if x.a.size == 1:
        return V(np.asfortranarray(np.eye(int(scalar(x)))))
    return V(np.asfortranarray(np.diag(as_dbl(x))))


def b_outer(it, pos, named):
    return V(np.asfortranarray(np.outer(as_dbl(pos[0]), as_dbl(pos[1]))))


def install(it):
    g = it.baseenv.vars

    def reg(name, fn, special=False):
        g[name] = Builtin(fn, name, special)
    for op in ("+", "-", "*", "/", "^", "%%", "%/%"):
        reg(op, arith(op))
    for op in ("==", "!=", "<", ">", "<=", ">="):
        reg(op, compare(op))
    reg("&", logic("&"))
    reg("|", logic("|"))
    reg("unary-", unary_minus)
    reg("unary+", unary_plus)
    reg("unary!", unary_not)
    reg("!", unary_not)
    reg(":", b_colon)
    reg("%*%", b_matmul)
    reg("%in%", b_in)
    reg("%o%", b_outer)
    for name, fn in (("log", np.log), ("exp", np.exp), ("sqrt", np.sqrt), ("digamma", sp.digamma),
                     ("lgamma", sp.gammaln), ("gamma", sp.gamma), ("floor", np.floor), ("ceiling", np.ceil),
                     ("lfactorial", lambda a: sp.gammaln(a + 1)), ("factorial", lambda a: sp.gamma(a + 1)),
                     ("log1p", np.log1p), ("expm1", np.expm1), ("log2", np.log2), ("log10", np.log10),
                     ("trigamma", lambda a: sp.polygamma(1, a)), ("sin", np.sin), ("cos", np.cos)):
        reg(name, math1(fn))
    reg("abs", math1(np.abs, keep_int=True))
    reg("round", b_round)
    for name, fn in (("c", b_c), ("list", b_list), ("length", b_length), ("dim", b_dim), ("nrow", b_nrow),
                     ("ncol", b_ncol), ("NROW", b_nrow), ("NCOL", b_ncol), ("seq", b_seq), ("seq_len", b_seq),
                     ("seq_along", lambda it_, p, n: intv(np.arange(1, (0 if p[0] is None else len(p[0])) + 1))),
                     ("rep", b_rep), ("matrix", b_matrix), ("t", b_t), ("crossprod", b_crossprod),
                     ("tcrossprod", b_tcrossprod), ("sum", b_sum), ("prod", b_prod), ("max", _minmax(np.max)),
                     ("min", _minmax(np.min)), ("mean", b_mean), ("var", b_var), ("median", b_median),
                     ("colSums", _colrow(0)), ("rowSums", _colrow(1)), ("colMeans", _colrow(0, True)),
                     ("rowMeans", _colrow(1, True)), ("cumsum", b_cumsum), ("any", b_any), ("all", b_all),
                     ("which", b_which), ("ifelse", b_ifelse), ("is.na", b_is_na), ("sweep", b_sweep),
                     ("apply", b_apply), ("lapply", b_lapply), ("sapply", b_sapply), ("unlist", b_unlist),
                     ("names", b_names), ("names<-", b_names_assign), ("rownames", _dimnames_get(0)),
                     ("colnames", _dimnames_get(1)), ("rownames<-", _dimnames_set(0)),
                     ("colnames<-", _dimnames_set(1)), ("setNames", b_setNames), ("dimnames", b_dimnames),
                     ("dimnames<-", b_dimnames_assign), ("class", b_class),
                     ("class<-", b_class_assign), ("inherits", b_inherits), ("attr", b_attr),
                     ("attr<-", b_attr_assign), ("as.vector", b_as_vector), ("as.numeric", b_as_numeric),
                     ("as.double", b_as_numeric), ("as.integer", b_as_integer), ("as.character", b_as_character),
                     ("as.logical", b_as_logical), ("as.matrix", b_as_matrix), ("all.equal", b_all_equal),
                     ("isTRUE", b_isTRUE), ("identical", b_identical), ("order", b_order), ("sort", b_sort),
                     ("rev", b_rev), ("unique", b_unique), ("duplicated", b_duplicated),
                     ("data.frame", b_data_frame), ("match", b_match), ("scale", b_scale), ("pnorm", b_pnorm),
                     ("qnorm", b_qnorm), ("paste0", b_paste("")), ("paste", b_paste(" ")), ("format", b_format),
                     ("cat", b_cat), ("print", b_print), ("stop", b_stop), ("warning", b_warning),
                     ("stopifnot", b_stopifnot), ("deparse", b_deparse), ("invisible", b_invisible),
                     ("do.call", b_do_call), ("file.path", b_file_path), ("diff", b_diff),
                     ("cbind", b_cbind(1)), ("rbind", b_cbind(0)), ("uniroot", b_uniroot), ("nchar", b_nchar),
                     ("numeric", b_numeric), ("diag", b_diag), ("outer", b_outer),
                     ("is.nan", b_is_elementwise(np.isnan)), ("is.finite", b_is_elementwise(np.isfinite)),
                     ("is.infinite", b_is_elementwise(np.isinf)), ("list.files", b_list_files),
                     ("file.remove", b_file_remove)):
        reg(name, fn)
    for kind in ("null", "list", "function", "vector", "matrix", "numeric", "double", "integer", "logical",
                 "character"):
        reg("is." + kind, b_is(kind))
    reg("dir.exists", lambda it_, p, n: lgl(os.path.isdir(p[0].a[0])))
    reg("file.exists", lambda it_, p, n: lgl([os.path.exists(f) for f in _strs(p[0])]))
    reg("Sys.time", lambda it_, p, n: dbl(0.0))
    for name, fn in (("with", sp_with), ("quote", sp_quote), ("substitute", sp_substitute), ("missing", sp_missing),
                     ("match.call", sp_match_call), ("rm", sp_rm), ("tryCatch", sp_tryCatch), ("return", sp_return),
                     ("library", sp_function_noop), ("require", sp_function_noop), ("set.seed", sp_function_noop),
                     ("suppressWarnings", sp_suppress), ("switch", sp_switch), ("on.exit", sp_function_noop),
                     ("save", sp_save)):
        reg(name, fn, special=True)
    g["pi"] = dbl(math.pi)
    g[".Machine"] = RList([dbl(np.finfo(np.float64).eps), intv(2 ** 31 - 1), dbl(np.finfo(np.float64).max),
                           dbl(np.finfo(np.float64).tiny)],
                          ["double.eps", "integer.max", "double.xmax", "double.xmin"])
    g["LETTERS"] = chr_(list("ABCDEFGHIJKLMNOPQRSTUVWXYZ"))
    g["letters"] = chr_(list("abcdefghijklmnopqrstuvwxyz"))
    it.namespaces["gsl"] = {"gamma_inc": Builtin(b_gamma_inc, "gsl::gamma_inc"),
                            "expint_E1": Builtin(b_expint_E1, "gsl::expint_E1"),
                            "hyperg_1F1": Builtin(b_hyperg_1F1, "gsl::hyperg_1F1")}
    it.namespaces["PowerTOST"] = {"OwensT": Builtin(b_OwensT, "PowerTOST::OwensT")}
