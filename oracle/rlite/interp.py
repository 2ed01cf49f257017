"""Evaluator for the R subset (see parser.py, values.py).  TEST INFRASTRUCTURE ONLY (part of oracle/).

Evaluation model: lexical scoping with environments, closures with lazily evaluated default arguments, eagerly
evaluated supplied arguments (their expressions are kept for substitute() / match.call() / missing()), replacement
calls (`f(x) <- v`, `x[i] <- v`, `x$a <- v`, nested), copy-on-assign for replacement (values are never modified in
place by R-level code), function lookup that skips non-function bindings (so a variable named `c` does not hide c()).
"""
from . import parser as P
from .values import (Builtin, Closure, Env, Lang, Promise, RError, RList, V, chr_, dbl, intv, lgl, truthy)


class ReturnEx(Exception):
    def __init__(self, value):
        self.value = value


class BreakEx(Exception):
    pass


class NextEx(Exception):
    pass


class Interp:
    def __init__(self, out=None):
        from . import base
        self.baseenv = Env(None)  # builtins; user definitions (and a variable named `c`) live one level below
        self.globalenv = Env(self.baseenv)
        self.warnings = []
        self.out = out if out is not None else []
        self.namespaces = {}
        base.install(self)

    # ------------------------------------------------------------------ entry points
    def source(self, path):
        src = open(path).read()
        return self.run(src, path)

    def run(self, src, filename="<text>", env=None):
        env = env or self.globalenv
        val = None
        for e in P.parse(src, filename):
            val = self.eval(e, env)
        return val

    def call(self, fname, *args, **kwargs):
        """Call an R function by name with already-evaluated R values."""
        fn = self.lookup_fn(fname, self.globalenv)
        return self.apply(fn, list(args), dict(kwargs), self.globalenv)

    def warn(self, msg):
        self.warnings.append(msg)

    # ------------------------------------------------------------------ variables
    def lookup(self, name, env):
        e = env
        while e is not None:
            if name in e.vars:
                v = e.vars[name]
                if isinstance(v, Promise):
                    v = self.force(v)
                    e.vars[name] = v
                return v
            e = e.parent
        raise RError(f"object '{name}' not found")

    def lookup_fn(self, name, env):
        e = env
        while e is not None:
            if name in e.vars:
                v = e.vars[name]
                if isinstance(v, Promise):
                    v = self.force(v)
                    e.vars[name] = v
                if isinstance(v, (Closure, Builtin)):
                    return v
            e = e.parent
        raise RError(f"could not find function \"{name}\"")

    def force(self, p):
        if not p.forced:
            p.value = self.eval(p.expr, p.env)
            p.forced = True
        return p.value

    # ------------------------------------------------------------------ eval
    def eval(self, e, env):
        k = e[0]
        if k == "id":
            return self.lookup(e[1], env)
        if k == "num":
            return intv(e[1]) if e[2] else dbl(e[1])
        if k == "str":
            return chr_(e[1])
        if k == "const":
            v = e[1]
            return None if v is None else (lgl(v) if isinstance(v, bool) else dbl(v))
        if k == "paren":
            return self.eval(e[1], env)
        if k == "block":
            val = None
            for x in e[1]:
                val = self.eval(x, env)
            return val
        if k == "call":
            return self.eval_call(e, env)
        if k == "binop":
            op = e[1]
            if op == "&&":
                if not truthy(self.eval(e[2], env), "&&"):
                    return lgl(False)
                return lgl(truthy(self.eval(e[3], env), "&&"))
            if op == "||":
                if truthy(self.eval(e[2], env), "||"):
                    return lgl(True)
                return lgl(truthy(self.eval(e[3], env), "||"))
            fn = self.lookup_fn(op, env)
            return fn.fn(self, [self.eval(e[2], env), self.eval(e[3], env)], {})
        if k == "unop":
            x = self.eval(e[2], env)
            return self.lookup_fn("unary" + e[1], env).fn(self, [x], {})
        if k == "assign":
            val = self.eval(e[2], env)
            self.assign(e[1], val, env, e[3])
            return val
        if k == "index":
            obj = self.eval(e[1], env)
            pos, named = self.eval_args(e[2], env, keep_empty=True)
            from . import base
            return base.index(self, obj, pos, named, e[3])
        if k == "dollar":
            obj = self.eval(e[1], env)
            from . import base
            return base.dollar(obj, e[2])
        if k == "ns":
            ns = self.namespaces.get(e[1], {})
            if e[2] in ns:
                return ns[e[2]]
            return self.lookup(e[2], env)
        if k == "function":
            return Closure(e[1], e[2], env)
        if k == "if":
            if truthy(self.eval(e[1], env), "if"):
                return self.eval(e[2], env)
            return self.eval(e[3], env) if e[3] is not None else None
        if k == "for":
            seq = self.eval(e[2], env)
            items = seq.items if isinstance(seq, RList) else ([] if seq is None else [V(seq.flat()[i:i + 1])
                                                                                      for i in range(seq.a.size)])
            for it in items:
                env.vars[e[1]] = it
                try:
                    self.eval(e[3], env)
                except BreakEx:
                    break
                except NextEx:
                    continue
            return None
        if k == "while":
            while truthy(self.eval(e[1], env), "while"):
                try:
                    self.eval(e[2], env)
                except BreakEx:
                    break
                except NextEx:
                    continue
            return None
        if k == "repeat":
            while True:
                try:
                    self.eval(e[1], env)
                except BreakEx:
                    break
                except NextEx:
                    continue
            return None
        if k == "break":
            raise BreakEx()
        if k == "next":
            raise NextEx()
        raise RError(f"cannot evaluate node {k}")

    def eval_args(self, args, env, keep_empty=False):
        pos, named = [], {}
        for name, ex in args:
            if ex is None:
                if keep_empty and name is None:
                    pos.append(Ellipsis)  # empty subscript
                continue
            if ex == ("id", "..."):
                dots = self.lookup("...", env)
                for n, v in dots:
                    if n:
                        named[n] = v
                    else:
                        pos.append(v)
                continue
            v = self.eval(ex, env)
            if name is None:
                pos.append(v)
            else:
                named[name] = v
        return pos, named

    def eval_call(self, e, env):
        fe = e[1]
        if fe[0] == "id":
            fn = self.lookup_fn(fe[1], env)
        elif fe[0] == "str":
            fn = self.lookup_fn(fe[1], env)
        else:
            fn = self.eval(fe, env)
        if isinstance(fn, Builtin) and fn.special:
            return fn.fn(self, env, e[2])
        pos, named = self.eval_args(e[2], env)
        return self.apply(fn, pos, named, env, call_expr=e)

    def apply(self, fn, pos, named, env, call_expr=None):
        if isinstance(fn, Builtin):
            if fn.special:
                raise RError(f"{fn.name} cannot be applied to evaluated arguments here")
            return fn.fn(self, pos, named)
        if not isinstance(fn, Closure):
            raise RError("attempt to apply non-function")
        fenv = Env(fn.env)
        fenv.call = call_expr
        fenv.fn = fn
        formals = [p[0] for p in fn.params]
        bound = {}
        named = dict(named)
        # 1. exact names, 2. unique partial matches (formals before ...), 3. positions
        for n in list(named):
            if n in formals and n != "...":
                bound[n] = named.pop(n)
        has_dots = "..." in formals
        before_dots = formals[:formals.index("...")] if has_dots else formals
        for n in list(named):
            cands = [f for f in before_dots if f.startswith(n) and f not in bound]
            if len(cands) == 1:
                bound[cands[0]] = named.pop(n)
        free = [f for f in formals if f not in bound and f != "..."]
        dots = []
        pos = list(pos)
        if has_dots:
            free_before = [f for f in before_dots if f not in bound]
            while pos and free_before:
                bound[free_before.pop(0)] = pos.pop(0)
            dots = [(None, v) for v in pos] + [(n, v) for n, v in named.items()]
            pos, named = [], {}
        else:
            while pos and free:
                bound[free.pop(0)] = pos.pop(0)
        if pos or named:
            raise RError(f"unused argument(s) in call to {fn.name or 'function'}: {len(pos)} positional, {list(named)}")
        for name, default in fn.params:
            if name == "...":
                fenv.vars["..."] = dots
            elif name in bound:
                fenv.vars[name] = bound[name]
            elif default is not None:
                fenv.vars[name] = Promise(default, fenv)
        fenv.vars[".supplied"] = set(bound)
        try:
            return self.eval(fn.body, fenv)
        except ReturnEx as r:
            return r.value

    # ------------------------------------------------------------------ assignment
    def assign(self, target, val, env, is_super=False):
        k = target[0]
        if k == "paren":
            return self.assign(target[1], val, env, is_super)
        if k in ("id", "str"):
            name = target[1]
            if is_super:
                e = env.parent
                while e is not None and name not in e.vars:
                    e = e.parent
                (e or self.globalenv).vars[name] = val
            else:
                env.vars[name] = val
            return
        from . import base
        if k == "index":
            obj = self.get_for_replace(target[1], env)
            pos, named = self.eval_args(target[2], env, keep_empty=True)
            new = base.index_assign(self, obj, pos, named, target[3], val)
            return self.assign(target[1], new, env, is_super)
        if k == "dollar":
            obj = self.get_for_replace(target[1], env)
            new = base.dollar_assign(obj, target[2], val)
            return self.assign(target[1], new, env, is_super)
        if k == "call":
            fname = target[1][1] + "<-"
            fn = self.lookup_fn(fname, env)
            inner = target[2][0][1]
            obj = self.get_for_replace(inner, env)
            pos, named = self.eval_args(target[2][1:], env)
            named["value"] = val
            new = self.apply(fn, [obj] + pos, named, env)
            return self.assign(inner, new, env, is_super)
        raise RError(f"invalid assignment target {P.deparse(target)}")

    def get_for_replace(self, expr, env):
        """Current value of a replacement target (NULL when a plain variable does not exist yet)."""
        if expr[0] == "id":
            try:
                return self.lookup(expr[1], env)
            except RError:
                return None
        return self.eval(expr, env)

    # ------------------------------------------------------------------ helpers for builtins
    def call_value(self, fn, pos, named=None):
        return self.apply(fn, pos, named or {}, self.globalenv)

    def match_call(self, env):
        if env.call is None:
            raise RError("match.call() outside a closure")
        exprs = [env.call[1]]
        for name, ex in env.call[2]:
            if ex == ("id", "..."):
                raise RError("match.call(): forwarding ... is not supported")
            exprs.append(ex)
        return Lang(exprs)
