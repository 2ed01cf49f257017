"""Tokeniser and Pratt parser for the subset of the R language the reference's R/*.R files use.

TEST INFRASTRUCTURE ONLY (part of oracle/).  The grammar follows the R language definition (operator table of
?Syntax); nothing here is derived from the reference package.  AST nodes are tuples:

    ("num", float|int, is_int) ("str", s) ("id", name) ("const", value)
    ("call", fn_expr, [(argname|None, expr|None), ...])     expr None = empty argument (x[, j])
    ("index", obj, args, double_bracket)  ("dollar", obj, name)  ("ns", pkg, name)
    ("binop", op, lhs, rhs) ("unop", op, e) ("assign", target, value, is_super)
    ("function", [(name, default|None)], body) ("if", c, a, b|None) ("for", var, seq, body)
    ("while", c, body) ("repeat", body) ("block", [expr]) ("break",) ("next",) ("paren", e)
"""
import re

TOKEN_RE = re.compile(r"""
    (?P<ws>[ \t\r\f]+)
  | (?P<comment>\#[^\n]*)
  | (?P<nl>\n)
  | (?P<num>(?:0[xX][0-9a-fA-F]+|(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?)L?)
  | (?P<str>"(?:[^"\\]|\\.)*"|'(?:[^'\\]|\\.)*')
  | (?P<bt>`[^`]+`)
  | (?P<id>(?:[A-Za-z]|\.(?![0-9]))[A-Za-z0-9._]*|\.)
  | (?P<op><<-|->>|<-|->|<=|>=|==|!=|&&|\|\||::|%[^%\n]*%|\[\[|[-+*/^<>=!&|~?:$@(){}\[\],;])
""", re.X)

ESCAPES = {"n": "\n", "t": "\t", "\\": "\\", '"': '"', "'": "'", "0": "\0", "r": "\r"}


class RSyntaxError(Exception):
    pass


def tokenize(src):
    toks, pos, line = [], 0, 1
    while pos < len(src):
        m = TOKEN_RE.match(src, pos)
        if not m:
            raise RSyntaxError(f"line {line}: cannot tokenise {src[pos:pos + 20]!r}")
        pos = m.end()
        k = m.lastgroup
        t = m.group()
        if k in ("ws", "comment"):
            continue
        if k == "nl":
            toks.append(("nl", "\n", line))
            line += 1
        elif k == "num":
            if t.endswith("L"):
                toks.append(("num", (int(t[:-1], 0) if t[:2].lower() == "0x" else int(float(t[:-1])), True), line))
            elif t[:2].lower() == "0x":
                toks.append(("num", (float(int(t, 16)), False), line))
            else:
                toks.append(("num", (float(t), False), line))
        elif k == "str":
            body = re.sub(r"\\(.)", lambda mm: ESCAPES.get(mm.group(1), mm.group(1)), t[1:-1])
            line += t.count("\n")
            toks.append(("str", body, line))
        elif k == "bt":
            toks.append(("id", t[1:-1], line))
        elif k == "id":
            toks.append(("id", t, line))
        else:
            toks.append(("op", t, line))
    toks.append(("eof", None, line))
    return toks


# binary operators: (left binding power, right binding power); right-assoc when rbp < lbp
BINOPS = {
    "?": (1, 2),
    "=": (4, 3), "<-": (6, 5), "<<-": (6, 5), "->": (7, 8), "->>": (7, 8),
    "~": (9, 10),
    "||": (11, 12), "|": (11, 12),
    "&&": (13, 14), "&": (13, 14),
    # "!" unary: 15
    "==": (17, 18), "!=": (17, 18), "<": (17, 18), ">": (17, 18), "<=": (17, 18), ">=": (17, 18),
    "+": (19, 20), "-": (19, 20),
    "*": (21, 22), "/": (21, 22),
    # %any% 23
    ":": (25, 26),
    # unary +/-: 27
    "^": (30, 29),
}
KEYWORD_CONST = {"TRUE": True, "FALSE": False, "T": True, "F": False, "NULL": None, "NA": float("nan"),
                 "NA_real_": float("nan"), "NA_integer_": float("nan"), "NA_character_": None,
                 "Inf": float("inf"), "NaN": float("nan")}
RESERVED = {"if", "else", "for", "while", "repeat", "function", "break", "next", "in"}


class Parser:
    def __init__(self, src, filename="<text>"):
        self.toks = tokenize(src)
        self.i = 0
        self.filename = filename
        self.depth = 0  # >0 inside ( or [ : newlines are not terminators there

    # -- token helpers
    def peek(self):
        if self.depth > 0:
            while self.toks[self.i][0] == "nl":
                self.i += 1
        return self.toks[self.i]

    def next(self):
        t = self.peek()
        self.i += 1
        return t

    def skip_nl(self):
        while self.toks[self.i][0] == "nl":
            self.i += 1

    def at_op(self, *ops):
        t = self.peek()
        return t[0] == "op" and t[1] in ops

    def expect_op(self, op):
        t = self.next()
        if t[0] != "op" or t[1] != op:
            raise RSyntaxError(f"{self.filename}:{t[2]}: expected {op!r}, got {t[1]!r}")
        return t

    def err(self, t, what):
        return RSyntaxError(f"{self.filename}:{t[2]}: {what} (token {t[1]!r})")

    # -- program
    def parse_program(self):
        out = []
        while True:
            self.skip_nl()
            while self.at_op(";"):
                self.next()
                self.skip_nl()
            if self.peek()[0] == "eof":
                return out
            out.append(self.parse_expr(0))
            t = self.toks[self.i]
            if t[0] not in ("nl", "eof") and not (t[0] == "op" and t[1] == ";"):
                raise self.err(t, "unexpected token after expression")

    # -- expressions
    def parse_expr(self, min_bp):
        lhs = self.parse_prefix()
        while True:
            t = self.peek()
            if t[0] == "op":
                op = t[1]
                # postfix forms bind tighter than everything
                if op == "(":
                    lhs = ("call", lhs, self.parse_args("(", ")"))
                    continue
                if op == "[[":
                    self.next()
                    self.depth += 1
                    args = self.parse_arglist_until("]")
                    self.expect_op("]")
                    self.depth -= 1
                    self.expect_op("]")
                    lhs = ("index", lhs, args, True)
                    continue
                if op == "[":
                    lhs = ("index", lhs, self.parse_args("[", "]"), False)
                    continue
                if op in ("$", "@"):
                    self.next()
                    n = self.next()
                    if n[0] not in ("id", "str"):
                        raise self.err(n, "name expected after $")
                    lhs = ("dollar", lhs, n[1])
                    continue
                if op == "::":
                    self.next()
                    n = self.next()
                    lhs = ("ns", lhs[1], n[1])
                    continue
                if op.startswith("%") and len(op) > 1:
                    lbp, rbp = 23, 24
                elif op in BINOPS:
                    lbp, rbp = BINOPS[op]
                else:
                    break
                if lbp < min_bp:
                    break
                self.next()
                self.skip_nl()
                rhs = self.parse_expr(rbp)
                if op in ("<-", "=", "<<-"):
                    lhs = ("assign", lhs, rhs, op == "<<-")
                elif op in ("->", "->>"):
                    lhs = ("assign", rhs, lhs, op == "->>")
                else:
                    lhs = ("binop", op, lhs, rhs)
                continue
            break
        return lhs

    def parse_prefix(self):
        t = self.next()
        kind, val = t[0], t[1]
        if kind == "num":
            return ("num", val[0], val[1])
        if kind == "str":
            return ("str", val)
        if kind == "id":
            if val in KEYWORD_CONST:
                return ("const", KEYWORD_CONST[val])
            if val == "function":
                return self.parse_function()
            if val == "if":
                return self.parse_if()
            if val == "for":
                self.expect_op("(")
                self.depth += 1
                var = self.next()
                kw = self.next()
                if kw[1] != "in":
                    raise self.err(kw, "'in' expected")
                seq = self.parse_expr(0)
                self.depth -= 1
                self.expect_op(")")
                return ("for", var[1], seq, self.parse_body())
            if val == "while":
                self.expect_op("(")
                self.depth += 1
                c = self.parse_expr(0)
                self.depth -= 1
                self.expect_op(")")
                return ("while", c, self.parse_body())
            if val == "repeat":
                return ("repeat", self.parse_body())
            if val == "break":
                return ("break",)
            if val == "next":
                return ("next",)
            return ("id", val)
        if kind == "op":
            if val == "(":
                self.depth += 1
                e = self.parse_expr(0)
                self.depth -= 1
                self.expect_op(")")
                return ("paren", e)
            if val == "{":
                return self.parse_block()
            if val in ("-", "+"):
                return ("unop", val, self.parse_expr(27))
            if val == "!":
                return ("unop", "!", self.parse_expr(15))
            if val == "~":
                return ("unop", "~", self.parse_expr(10))
        raise self.err(t, "unexpected token")

    def parse_body(self):
        self.skip_nl()
        return self.parse_expr(0)

    def parse_block(self):
        saved, self.depth = self.depth, 0
        exprs = []
        while True:
            self.skip_nl()
            while self.at_op(";"):
                self.next()
                self.skip_nl()
            if self.at_op("}"):
                self.next()
                break
            exprs.append(self.parse_expr(0))
            t = self.toks[self.i]
            if not (t[0] == "nl" or (t[0] == "op" and t[1] in (";", "}"))):
                raise self.err(t, "unexpected token in block")
        self.depth = saved
        return ("block", exprs)

    def parse_if(self):
        self.expect_op("(")
        self.depth += 1
        c = self.parse_expr(0)
        self.depth -= 1
        self.expect_op(")")
        a = self.parse_body()
        # `else` may follow on a later line (legal inside braces, which is where the sources use it)
        j = self.i
        while self.toks[j][0] == "nl":
            j += 1
        b = None
        if self.toks[j][0] == "id" and self.toks[j][1] == "else":
            self.i = j + 1
            b = self.parse_body()
        return ("if", c, a, b)

    def parse_function(self):
        self.expect_op("(")
        self.depth += 1
        params = []
        while not self.at_op(")"):
            n = self.next()
            if n[0] != "id":
                raise self.err(n, "formal argument name expected")
            default = None
            if self.at_op("="):
                self.next()
                default = self.parse_expr(5)
            params.append((n[1], default))
            if self.at_op(","):
                self.next()
        self.depth -= 1
        self.expect_op(")")
        return ("function", params, self.parse_body())

    def parse_args(self, open_, close):
        self.expect_op(open_)
        self.depth += 1
        args = self.parse_arglist_until(close)
        self.depth -= 1
        self.expect_op(close)
        return args

    def parse_arglist_until(self, close):
        args = []
        if self.at_op(close):
            return args
        while True:
            if self.at_op(",") or self.at_op(close):
                args.append((None, None))  # empty argument
            else:
                t = self.peek()
                # name = value (the name may be a symbol or a string)
                j = self.i + 1
                while self.toks[j][0] == "nl":
                    j += 1
                nxt = self.toks[j]
                if t[0] in ("id", "str") and nxt[0] == "op" and nxt[1] == "=" and t[1] not in RESERVED:
                    self.next()
                    self.next()
                    if self.at_op(",") or self.at_op(close):
                        args.append((t[1], None))
                    else:
                        args.append((t[1], self.parse_expr(5)))
                else:
                    args.append((None, self.parse_expr(5)))
            if self.at_op(","):
                self.next()
                if self.at_op(close):
                    args.append((None, None))
                    break
                continue
            break
        return args


def parse(src, filename="<text>"):
    return Parser(src, filename).parse_program()


def deparse(e):
    """Source text of an expression (enough for symbols, constants, calls and operators)."""
    k = e[0]
    if k == "id":
        return e[1]
    if k == "num":
        v = e[1]
        return (str(int(v)) + ("L" if e[2] else "")) if float(v).is_integer() else repr(v)
    if k == "str":
        return '"' + e[1] + '"'
    if k == "const":
        v = e[1]
        return "NULL" if v is None else ("TRUE" if v is True else "FALSE" if v is False else
                                         "NA" if v != v else "Inf")
    if k == "call":
        return deparse(e[1]) + "(" + ", ".join((f"{n} = " if n else "") + (deparse(a) if a else "")
                                               for n, a in e[2]) + ")"
    if k == "index":
        o, c = ("[[", "]]") if e[3] else ("[", "]")
        return deparse(e[1]) + o + ", ".join((f"{n} = " if n else "") + (deparse(a) if a else "")
                                             for n, a in e[2]) + c
    if k == "dollar":
        return deparse(e[1]) + "$" + e[2]
    if k == "ns":
        return e[1] + "::" + e[2]
    if k == "binop":
        return f"{deparse(e[2])} {e[1]} {deparse(e[3])}"
    if k == "unop":
        return e[1] + deparse(e[2])
    if k == "paren":
        return "(" + deparse(e[1]) + ")"
    return f"<{k}>"
