"""f90run — a small interpreter for the Fortran subset used by the uDALES hot-path routines.

Purpose (TEST INFRASTRUCTURE): the reference cannot be compiled in this image (no Fortran compiler,
MPI, FFTW), and it ships no golden vectors for its numerical routines.  To pin the CPU oracle to the
reference anyway, this interpreter executes the reference's OWN SOURCE TEXT (read from
/root/reference at generation time, never copied into this repository) on seeded inputs, in IEEE
double arithmetic with Fortran evaluation order, and `make_golden.py` stores inputs + outputs as
tests/golden/*.npz.  tests/test_oracle_golden.py then checks oracle/udales_oracle.c against them.

Supported: free-form source, `&` continuations, comments; subroutines / functions inside modules;
declarations with explicit-shape, assumed/deferred shape and automatic arrays; do / do-while-free
loops, if / else if / else, one-line if, select case, call, return, allocate / deallocate, whole-array
and array-section assignment, array constructors, keyword arguments; the intrinsics the hot path
uses.  `use`, `implicit`, `external`, I/O statements are ignored.  External procedures (MPI,
2decomp, FFTW) are Python callbacks supplied by the harness.

Numerics: default real = IEEE double (-fdefault-real-8, CMakeLists.txt:46 of the reference); integer
division truncates; `x**n` with integer (or integral real) n is evaluated by repeated multiplication
as gfortran does; everything else follows the source's own operator order.
"""
from __future__ import annotations

import math
import re

import numpy as np


# ----------------------------------------------------------------------------------------------
class FArray:
    """numpy array in Fortran order + lower bounds."""

    __slots__ = ("a", "lb")

    def __init__(self, a, lb):
        self.a = a
        self.lb = tuple(int(x) for x in lb)

    @staticmethod
    def alloc(bounds, dtype=float, fill=np.nan):
        shape = tuple(max(0, hi - lo + 1) for lo, hi in bounds)
        a = np.empty(shape, dtype=dtype, order="F")
        a[...] = fill if dtype != int else 0
        return FArray(a, [lo for lo, _ in bounds])

    def rebound(self, bounds):
        """explicit-shape dummy: same storage, new bounds (sequence association)."""
        shape = tuple(hi - lo + 1 for lo, hi in bounds)
        if int(np.prod(shape)) > self.a.size:
            raise ValueError(f"dummy larger than actual: {shape} vs {self.a.shape}")
        if shape == self.a.shape:
            return FArray(self.a, [lo for lo, _ in bounds])
        flat = self.a.reshape(-1, order="F")
        if not np.shares_memory(flat, self.a):
            raise ValueError("non-contiguous actual argument for an explicit-shape dummy")
        return FArray(flat[:int(np.prod(shape))].reshape(shape, order="F"), [lo for lo, _ in bounds])


class Slice:
    __slots__ = ("lo", "hi", "st")

    def __init__(self, lo, hi, st=None):
        self.lo, self.hi, self.st = lo, hi, st


class ReturnSignal(Exception):
    pass


class StopSignal(Exception):
    pass


# ----------------------------------------------------------------------------------------------
TOKEN_RE = re.compile(r"""
    (?P<num>(\d+\.\d*|\.\d+|\d+)([edED][+-]?\d+)?(_\w+)?)
  | (?P<dotop>\.(and|or|not|eqv|neqv|eq|ne|lt|le|gt|ge|true|false)\.)
  | (?P<name>[A-Za-z_]\w*)
  | (?P<str>'[^']*'|"[^"]*")
  | (?P<op>\*\*|==|/=|<=|>=|=>|\(/|/\)|::|[-+*/(),=<>:%])
  | (?P<ws>\s+)
""", re.X | re.I)


def tokenize(s):
    out = []
    pos = 0
    while pos < len(s):
        m = TOKEN_RE.match(s, pos)
        if not m:
            raise SyntaxError(f"cannot tokenize: {s[pos:pos + 30]!r} in {s!r}")
        pos = m.end()
        k = m.lastgroup
        if k == "ws":
            continue
        t = m.group(k)
        if k == "num":
            t = re.sub(r"_\w+$", "", t)
            isreal = bool(re.search(r"[.edED]", t))
            out.append(("num", float(t.lower().replace("d", "e")) if isreal else int(t)))
        elif k == "dotop":
            out.append(("op", t.lower()))
        elif k == "name":
            out.append(("name", t.lower()))
        elif k == "str":
            out.append(("str", t[1:-1]))
        else:
            out.append(("op", t))
    # derived-type component reference a%b -> one name token "a%b" (resolved through a dict, see Interp.lookup)
    merged = []
    i = 0
    while i < len(out):
        if (out[i][0] == "name" and i + 2 < len(out) and out[i + 1] == ("op", "%") and out[i + 2][0] == "name"):
            merged.append(("name", out[i][1] + "%" + out[i + 2][1]))
            i += 3
        else:
            merged.append(out[i])
            i += 1
    return merged


class Parser:
    """expression parser -> tuple AST"""

    def __init__(self, toks):
        self.t = toks
        self.i = 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else ("eof", None)

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def accept(self, v):
        if self.peek() == ("op", v):
            self.i += 1
            return True
        return False

    def expect(self, v):
        if not self.accept(v):
            raise SyntaxError(f"expected {v!r} at {self.t[self.i:self.i + 5]}")

    # precedence: .or. < .and. < .not. < relational < +,- < *,/ < unary < **
    def expr(self):
        e = self.and_()
        while self.peek() in (("op", ".or."), ("op", ".eqv."), ("op", ".neqv.")):
            op = self.next()[1]
            e = ("bin", op, e, self.and_())
        return e

    def and_(self):
        e = self.not_()
        while self.accept(".and."):
            e = ("bin", ".and.", e, self.not_())
        return e

    def not_(self):
        if self.accept(".not."):
            return ("not", self.not_())
        return self.rel()

    REL = {"==": "==", ".eq.": "==", "/=": "/=", ".ne.": "/=", "<": "<", ".lt.": "<", "<=": "<=", ".le.": "<=",
           ">": ">", ".gt.": ">", ">=": ">=", ".ge.": ">="}

    def rel(self):
        e = self.add()
        tok = self.peek()
        if tok[0] == "op" and tok[1] in self.REL:
            self.next()
            e = ("bin", self.REL[tok[1]], e, self.add())
        return e

    def add(self):
        if self.accept("-"):
            e = ("neg", self.mul())
        elif self.accept("+"):
            e = self.mul()
        else:
            e = self.mul()
        while self.peek() in (("op", "+"), ("op", "-")):
            op = self.next()[1]
            e = ("bin", op, e, self.mul())
        return e

    def mul(self):
        e = self.pow_()
        while self.peek() in (("op", "*"), ("op", "/")):
            op = self.next()[1]
            e = ("bin", op, e, self.pow_())
        return e

    def pow_(self):
        base = self.unary_atom()
        if self.accept("**"):
            # right associative; exponent may carry a unary minus
            if self.accept("-"):
                ex = ("neg", self.pow_())
            else:
                ex = self.pow_()
            return ("bin", "**", base, ex)
        return base

    def unary_atom(self):
        if self.accept("-"):
            return ("neg", self.unary_atom())
        if self.accept("+"):
            return self.unary_atom()
        return self.atom()

    def atom(self):
        k, v = self.next()
        if k == "num":
            return ("num", v)
        if k == "str":
            return ("str", v)
        if (k, v) == ("op", ".true."):
            return ("num", True)
        if (k, v) == ("op", ".false."):
            return ("num", False)
        if (k, v) == ("op", "("):
            e = self.expr()
            if self.accept(","):               # complex literal (a, b)
                im = self.expr()
                self.expect(")")
                return ("call", "cmplx", [e, im], {})
            self.expect(")")
            return ("paren", e)
        if (k, v) == ("op", "(/"):
            items = []
            if not self.accept("/)"):
                items.append(self.expr())
                while self.accept(","):
                    items.append(self.expr())
                self.expect("/)")
            return ("arrcon", items)
        if k == "name":
            node = ("var", v)
            while True:
                if self.accept("("):
                    args, kw = self.arglist()
                    node = ("call", node[1], args, kw) if node[0] == "var" else ("index", node, args)
                elif self.accept("%"):
                    comp = self.next()[1]
                    node = ("comp", node, comp)
                else:
                    break
            return node
        raise SyntaxError(f"unexpected token {k, v} in {self.t}")

    def arglist(self):
        args, kw = [], {}
        if self.accept(")"):
            return args, kw
        while True:
            # keyword argument?
            if self.peek()[0] == "name" and self.i + 1 < len(self.t) and self.t[self.i + 1] == ("op", "=") :
                name = self.next()[1]
                self.next()
                kw[name] = self.expr()
            else:
                args.append(self.subscript())
            if self.accept(")"):
                break
            self.expect(",")
        return args, kw

    def subscript(self):
        lo = hi = st = None
        if self.peek() != ("op", ":"):
            lo = self.expr()
            if self.peek() != ("op", ":"):
                return lo
        self.expect(":")
        if self.peek() not in (("op", ","), ("op", ")"), ("op", ":")):
            hi = self.expr()
        if self.accept(":"):
            st = self.expr()
        return ("slice", lo, hi, st)


def parse_expr(s):
    p = Parser(tokenize(s))
    e = p.expr()
    if p.peek()[0] != "eof":
        raise SyntaxError(f"trailing tokens in expression {s!r}: {p.t[p.i:]}")
    return e


# ----------------------------------------------------------------------------------------------
def logical_lines(text):
    """strip comments, join continuations, split on ';'; yields (lineno, statement)."""
    out = []
    cur, cur_no = "", 0
    for no, raw in enumerate(text.split("\n"), 1):
        # remove comments (respecting quotes)
        line, q = "", None
        for ch in raw:
            if q:
                line += ch
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
                line += ch
            elif ch == "!":
                break
            else:
                line += ch
        line = line.strip()
        if not line:
            continue
        if line.startswith("#"):
            continue
        if line.startswith("&"):
            line = line[1:].lstrip()
        if not cur:
            cur_no = no
        if line.endswith("&"):
            cur += line[:-1] + " "
            continue
        cur += line
        for part in split_semicolons(cur):
            if part.strip():
                out.append((cur_no, part.strip()))
        cur = ""
    return out


def split_semicolons(s):
    parts, cur, q = [], "", None
    for ch in s:
        if q:
            cur += ch
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur += ch
        elif ch == ";":
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return parts


def split_top(s, sep=","):
    """split at top-level separators (outside parentheses / quotes)."""
    parts, cur, depth, q = [], "", 0, None
    i = 0
    while i < len(s):
        ch = s[i]
        if q:
            cur += ch
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur += ch
        elif ch == "(":
            depth += 1
            cur += ch
        elif ch == ")":
            depth -= 1
            cur += ch
        elif ch == sep and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
        i += 1
    parts.append(cur)
    return [p.strip() for p in parts]


def match_paren(s, start):
    depth = 0
    for i in range(start, len(s)):
        if s[i] == "(":
            depth += 1
        elif s[i] == ")":
            depth -= 1
            if depth == 0:
                return i
    raise SyntaxError(f"unbalanced parentheses in {s!r}")


DECL_RE = re.compile(r"^(real|integer|logical|complex|character|double\s+precision|type\s*\()", re.I)
IGNORE_RE = re.compile(r"^(use\b|implicit\b|external\b|include\b|save\b|private\b|public\b|intent\b|parameter\b|"
                       r"write\b|print\b|open\b|close\b|read\b|format\b|\d+\s+format\b|namelist\b|contains\b|data\b|interface\b)", re.I)


class Proc:
    def __init__(self, name, args, kind, body, result=None, src=""):
        self.name, self.args, self.kind, self.body, self.result, self.src = name, args, kind, body, result, src
        self.decls = []      # (name, typ, dims or None, init expr or None, allocatable)
        self.code = None


def parse_decl(stmt):
    """returns list of (name, typ, dims, init, allocatable)."""
    low = stmt
    m = re.match(r"^(double\s+precision|type\s*\([^)]*\)|[a-z]+)\s*(\([^)]*\))?", low, re.I)
    typ = m.group(1).lower()
    rest = low[m.end():]
    attrs, ents = "", rest
    if "::" in rest:
        attrs, ents = rest.split("::", 1)
    attr_list = split_top(attrs.strip().lstrip(","), ",") if attrs.strip() else []
    dims_attr, alloc = None, False
    for a in attr_list:
        al = a.lower().strip()
        if al.startswith("dimension"):
            dims_attr = a[a.index("(") + 1:match_paren(a, a.index("("))]
        if al == "allocatable":
            alloc = True
    out = []
    for ent in split_top(ents, ","):
        if not ent:
            continue
        init = None
        if "=" in ent and "=>" not in ent:
            # careful: '=' inside parentheses does not occur in entity decls of the subset
            ent, init = ent.split("=", 1)
            ent, init = ent.strip(), init.strip()
        mm = re.match(r"^([A-Za-z_]\w*)\s*(\(.*\))?\s*$", ent.strip())
        if not mm:
            raise SyntaxError(f"cannot parse entity {ent!r} in {stmt!r}")
        name = mm.group(1).lower()
        dims = mm.group(2)[1:-1] if mm.group(2) else dims_attr
        out.append((name, typ, dims, init, alloc))
    return out


# ----------------------------------------------------------------------------------------------
class Interp:
    def __init__(self, g=None, externals=None, trace=False):
        self.g = g if g is not None else {}      # module-level variables (flat namespace, lower case)
        self.procs = {}
        self.ext = externals or {}
        self.trace = trace
        self.gtypes = {}        # module-level arrays that are not real (e.g. complex FFT buffers)
        self.intr = self._intrinsics()

    # -------- loading ---------------------------------------------------------------------
    def load(self, path, only=None):
        text = open(path, errors="ignore").read()
        lines = logical_lines(text)
        i = 0
        while i < len(lines):
            no, st = lines[i]
            m = re.match(r"^(?:(?:pure|elemental|recursive)\s+)*(?:(real|integer|logical)\s+)?(subroutine|function)\s+(\w+)\s*(\(([^)]*)\))?"
                         r"(?:\s*result\s*\(\s*(\w+)\s*\))?", st, re.I)
            if m and not re.match(r"^end\b", st, re.I):
                kind = m.group(2).lower()
                name = m.group(3).lower()
                args = [a.strip().lower() for a in (m.group(5) or "").split(",") if a.strip()]
                res = (m.group(6) or name).lower() if kind == "function" else None
                j = i + 1
                body = []
                while j < len(lines) and not re.match(rf"^end\s*(subroutine|function)?(\s+{name})?\s*$", lines[j][1], re.I):
                    body.append(lines[j])
                    j += 1
                if only is None or name in only:
                    p = Proc(name, args, kind, body, res, path)
                    if kind == "function" and m.group(1):
                        p.decls.append((res, m.group(1).lower(), None, None, False))
                    self.procs[name] = p
                i = j + 1
            else:
                i += 1

    # -------- compile body into block AST -----------------------------------------------------
    def compile(self, p: Proc):
        if p.code is not None:
            return
        stmts = []
        for no, st in p.body:
            if IGNORE_RE.match(st):
                continue
            if DECL_RE.match(st) and ("::" in st or re.match(r"^(real|integer|logical|complex)\s+[A-Za-z_]", st, re.I)) \
                    and not re.match(r"^real\s*\(", st, re.I):
                p.decls.extend(parse_decl(st))
                continue
            stmts.append((no, st))
        self._pos = 0
        self._stmts = stmts
        p.code = self._block(terminators=())

    def _block(self, terminators):
        out = []
        while self._pos < len(self._stmts):
            no, st = self._stmts[self._pos]
            low = st.lower()
            head = re.match(r"^[a-z_]\w*", low)
            word = head.group(0) if head else ""
            if any(re.match(t, low) for t in terminators):
                return out
            self._pos += 1
            try:
                if re.match(r"^do\b", low) and not re.match(r"^do\s*while", low):
                    m = re.match(r"^do\s+(\w+)\s*=\s*(.*)$", st, re.I)
                    parts = split_top(m.group(2), ",")
                    body = self._block((r"^end\s*do\b", r"^enddo\b"))
                    self._pos += 1
                    out.append(("do", m.group(1).lower(), parse_expr(parts[0]), parse_expr(parts[1]),
                                parse_expr(parts[2]) if len(parts) > 2 else None, body, no))
                elif re.match(r"^if\s*\(", low):
                    close = match_paren(st, st.index("("))
                    cond = parse_expr(st[st.index("(") + 1:close])
                    tail = st[close + 1:].strip()
                    if tail.lower() == "then":
                        branches = []
                        body = self._block((r"^else\b", r"^elseif\b", r"^end\s*if\b", r"^endif\b"))
                        branches.append((cond, body))
                        while True:
                            no2, st2 = self._stmts[self._pos]
                            l2 = st2.lower()
                            self._pos += 1
                            if re.match(r"^(else\s*if|elseif)\s*\(", l2):
                                c2 = match_paren(st2, st2.index("("))
                                cnd = parse_expr(st2[st2.index("(") + 1:c2])
                                body = self._block((r"^else\b", r"^elseif\b", r"^end\s*if\b", r"^endif\b"))
                                branches.append((cnd, body))
                            elif re.match(r"^else\b", l2):
                                body = self._block((r"^end\s*if\b", r"^endif\b"))
                                branches.append((None, body))
                            else:
                                break
                        out.append(("if", branches, no))
                    else:
                        out.append(("if", [(cond, [self._simple(tail, no)])], no))
                elif re.match(r"^select\s*case", low):
                    close = match_paren(st, st.index("("))
                    sel = parse_expr(st[st.index("(") + 1:close])
                    cases = []
                    # skip to first case
                    while True:
                        no2, st2 = self._stmts[self._pos]
                        l2 = st2.lower()
                        if re.match(r"^end\s*select", l2):
                            self._pos += 1
                            break
                        if re.match(r"^case\s*default", l2):
                            self._pos += 1
                            body = self._block((r"^case\b", r"^end\s*select"))
                            cases.append((None, body))
                        elif re.match(r"^case\s*\(", l2):
                            self._pos += 1
                            c2 = match_paren(st2, st2.index("("))
                            vals = [parse_expr(v) for v in split_top(st2[st2.index("(") + 1:c2], ",")]
                            body = self._block((r"^case\b", r"^end\s*select"))
                            cases.append((vals, body))
                        else:
                            raise SyntaxError(f"unexpected in select: {st2}")
                    out.append(("select", sel, cases, no))
                else:
                    out.append(self._simple(st, no))
            except SyntaxError as e:
                raise SyntaxError(f"line {no}: {st!r}: {e}") from None
        return out

    def _simple(self, st, no):
        low = st.lower().strip()
        if low.startswith("call "):
            m = re.match(r"^call\s+(\w+)\s*(\(.*\))?\s*$", st, re.I | re.S)
            name = m.group(1).lower()
            args, kw = [], {}
            if m.group(2):
                p = Parser(tokenize(m.group(2)))
                p.expect("(")
                args, kw = p.arglist()
            return ("callst", name, args, kw, no)
        if low == "return":
            return ("return", no)
        if low.startswith("stop"):
            return ("stop", no)
        if low.startswith("allocate"):
            inner = st[st.index("(") + 1:match_paren(st, st.index("("))]
            items = []
            for it in split_top(inner, ","):
                if re.match(r"^\s*stat\s*=", it, re.I):
                    continue
                mm = re.match(r"^(\w+)\s*\((.*)\)$", it.strip(), re.S)
                items.append((mm.group(1).lower(), mm.group(2)))
            return ("allocate", items, no)
        if low.startswith("deallocate"):
            return ("nop", no)
        if low in ("continue", "cycle", "exit"):
            return (low, no)
        # assignment: find top-level '=' that is not part of ==, /=, <=, >=
        depth = 0
        for i, ch in enumerate(st):
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "=" and depth == 0:
                if st[i + 1:i + 2] == "=" or st[i - 1] in "=/<>":
                    continue
                lhs, rhs = st[:i].strip(), st[i + 1:].strip()
                return ("assign", parse_expr(lhs), parse_expr(rhs), no)
        raise SyntaxError(f"cannot parse statement {st!r}")

    # -------- evaluation ---------------------------------------------------------------------
    def _intrinsics(self):
        def fmax(*a):
            r = a[0]
            for x in a[1:]:
                r = np.maximum(r, x) if isinstance(r, np.ndarray) or isinstance(x, np.ndarray) else (r if r >= x else x)
            return r

        def fmin(*a):
            r = a[0]
            for x in a[1:]:
                r = np.minimum(r, x) if isinstance(r, np.ndarray) or isinstance(x, np.ndarray) else (r if r <= x else x)
            return r

        def freal(x, *k):
            if isinstance(x, np.ndarray):
                return x.real.astype(float)
            return float(x.real) if isinstance(x, complex) else float(x)

        def fint(x, *k):
            return int(x)          # truncation toward zero

        def fmod(a, b):
            if isinstance(a, int) and isinstance(b, int):
                return int(math.fmod(a, b))
            return math.fmod(a, b)

        def fsqrt(x):
            return np.sqrt(x) if isinstance(x, np.ndarray) else math.sqrt(x)

        def fabs_(x):
            return np.abs(x) if isinstance(x, np.ndarray) else abs(x)

        def fsum(x, *a):
            return float(np.sum(x)) if not np.iscomplexobj(x) else complex(np.sum(x))

        return {
            "sqrt": fsqrt, "abs": fabs_, "max": fmax, "min": fmin, "mod": fmod, "real": freal, "float": freal, "dble": freal,
            "int": fint, "nint": lambda x: int(round(x)), "sin": math.sin, "cos": math.cos, "exp": math.exp, "log": math.log,
            "atan": math.atan, "tanh": math.tanh,
            "cmplx": lambda a, b=0.0, *k: complex(a, b), "aimag": lambda z: (z.imag if not isinstance(z, np.ndarray) else z.imag.astype(float)),
            "sum": fsum, "kind": lambda x: 8, "sign": lambda a, b: math.copysign(abs(a), b),
            "maxval": lambda x: float(np.max(x)), "minval": lambda x: float(np.min(x)),
        }

    def call(self, name, *args, **kwargs):
        """call an interpreted procedure from Python (args: python scalars / FArray)."""
        return self._invoke(name.lower(), list(args), kwargs, None)

    def _invoke(self, name, argvals, kw, setters):
        p = self.procs[name]
        self.compile(p)
        loc = {}
        for an, av in zip(p.args, argvals):
            loc[an] = av
        for k, v in kw.items():
            loc[k] = v
        frame = {"loc": loc, "proc": p}
        # declarations: dummies get their declared bounds, automatic arrays are allocated
        for (nm, typ, dims, init, alloc) in p.decls:
            if nm in p.args:
                if dims is not None and nm in loc and isinstance(loc[nm], FArray):
                    parts = split_top(dims, ",")
                    if all(":" != d.strip() and d.strip() != "*" for d in parts) and not any(d.strip().endswith(":") for d in parts):
                        b = [self._bounds(d, frame) for d in parts]
                        loc[nm] = loc[nm].rebound(b)
                    elif any(d.strip().endswith(":") and d.strip() != ":" for d in parts):
                        lbs = [int(self.eval(parse_expr(d.strip()[:-1]), frame)) if d.strip() != ":" else 1 for d in parts]
                        loc[nm] = FArray(loc[nm].a, lbs)
                    else:
                        loc[nm] = FArray(loc[nm].a, [1] * loc[nm].a.ndim)   # assumed shape: lower bound 1
                continue
            if dims is not None and not alloc and ":" not in [d.strip() for d in split_top(dims, ",")]:
                b = [self._bounds(d, frame) for d in split_top(dims, ",")]
                dt = complex if typ == "complex" else int if typ == "integer" else float
                loc[nm] = FArray.alloc(b, dt)
            elif dims is None:
                if init is not None:
                    loc[nm] = self.eval(parse_expr(init), frame)
                elif nm not in loc:
                    loc[nm] = None
            else:
                loc[nm] = None     # allocatable, unallocated
        try:
            self.exec_block(p.code, frame)
        except ReturnSignal:
            pass
        # copy-out of scalar dummies (by reference semantics)
        if setters:
            for an, st in zip(p.args, setters):
                if st is not None and not isinstance(loc.get(an), FArray):
                    st(loc.get(an))
        if p.kind == "function":
            return loc.get(p.result)
        return None

    def _bounds(self, d, frame):
        d = d.strip()
        parts = split_top(d, ":")
        if len(parts) == 1:
            return (1, int(self.eval(parse_expr(parts[0]), frame)))
        return (int(self.eval(parse_expr(parts[0]), frame)), int(self.eval(parse_expr(parts[1]), frame)))

    def lookup(self, name, frame):
        loc = frame["loc"]
        if "%" in name:
            base, comp = name.split("%", 1)
            return self.lookup(base, frame)[comp]
        if name in loc:
            return loc[name]
        if name in self.g:
            return self.g[name]
        raise NameError(f"undefined variable {name!r} in {frame['proc'].name}")

    def store(self, name, val, frame):
        loc = frame["loc"]
        if "%" in name:
            base, comp = name.split("%", 1)
            self.lookup(base, frame)[comp] = val
            return
        if name in loc or name not in self.g:
            if name not in loc and name not in self.g:
                raise NameError(f"assignment to undeclared {name!r} in {frame['proc'].name}")
            loc[name] = val
        else:
            self.g[name] = val

    def exec_block(self, block, frame):
        for st in block:
            k = st[0]
            if k == "assign":
                self.assign(st[1], self.eval(st[2], frame), frame)
            elif k == "do":
                _, var, lo, hi, step, body, no = st
                a, b = int(self.eval(lo, frame)), int(self.eval(hi, frame))
                s = int(self.eval(step, frame)) if step is not None else 1
                i = a
                if s > 0:
                    while i <= b:
                        self.store(var, i, frame)
                        self.exec_block(body, frame)
                        i += s
                else:
                    while i >= b:
                        self.store(var, i, frame)
                        self.exec_block(body, frame)
                        i += s
                self.store(var, i, frame)
            elif k == "if":
                for cond, body in st[1]:
                    if cond is None or self.truth(self.eval(cond, frame)):
                        self.exec_block(body, frame)
                        break
            elif k == "select":
                v = self.eval(st[1], frame)
                done = False
                default = None
                for vals, body in st[2]:
                    if vals is None:
                        default = body
                        continue
                    if any(self.eval(x, frame) == v for x in vals):
                        self.exec_block(body, frame)
                        done = True
                        break
                if not done and default is not None:
                    self.exec_block(default, frame)
            elif k == "callst":
                self.do_call(st[1], st[2], st[3], frame)
            elif k == "return":
                raise ReturnSignal()
            elif k == "stop":
                raise StopSignal(f"STOP at line {st[1]} of {frame['proc'].name}")
            elif k == "allocate":
                for nm, dims in st[1]:
                    b = [self._bounds(d, frame) for d in split_top(dims, ",")]
                    typ = next((t for (n2, t, _, _, _) in frame["proc"].decls if n2 == nm), "real")
                    dt = complex if typ == "complex" else int if typ == "integer" else float
                    if nm in self.gtypes and nm not in frame["loc"]:
                        dt = self.gtypes[nm]
                    arr = FArray.alloc(b, dt)
                    if nm in frame["loc"] or nm not in self.g:
                        frame["loc"][nm] = arr
                    else:
                        self.g[nm] = arr
            elif k in ("nop", "continue"):
                pass
            else:
                raise NotImplementedError(k)

    @staticmethod
    def truth(v):
        if isinstance(v, np.ndarray):
            raise ValueError("array-valued condition")
        return bool(v)

    def do_call(self, name, args, kw, frame):
        if name in self.procs:
            vals, setters = [], []
            for a in args:
                vals.append(self.eval_arg(a, frame))
                setters.append(self.make_setter(a, frame))
            kwv = {k: self.eval_arg(v, frame) for k, v in kw.items()}
            self._invoke(name, vals, kwv, setters)
        elif name in self.ext:
            vals = [self.eval_arg(a, frame) for a in args]
            setters = [self.make_setter(a, frame) for a in args]
            kwv = {k: self.eval_arg(v, frame) for k, v in kw.items()}
            self.ext[name](self, frame, vals, setters, kwv)
        else:
            raise NameError(f"call to unknown procedure {name!r} in {frame['proc'].name}")

    def make_setter(self, node, frame):
        if node[0] == "var":
            nm = node[1]
            return lambda v, nm=nm: self.store(nm, v, frame)
        if node[0] == "call" and self.is_array(node[1], frame):
            return lambda v, node=node: self.assign(node, v, frame)
        return None

    def is_array(self, name, frame):
        try:
            return isinstance(self.lookup(name, frame), FArray)
        except NameError:
            return False

    def eval_arg(self, node, frame):
        """actual argument: arrays / sections are passed by reference (views)."""
        if node[0] == "var":
            return self.lookup(node[1], frame)
        if node[0] == "call" and self.is_array(node[1], frame):
            arr = self.lookup(node[1], frame)
            idx, is_sec = self.index_of(arr, node[2], frame)
            if is_sec:
                view = arr.a[idx]
                return FArray(view, [1] * view.ndim)
            # scalar element passed where array dummy expected: sequence association from that element
            return arr.a[idx]
        return self.eval(node, frame)

    def index_of(self, arr, subs, frame):
        idx, is_sec = [], False
        for d, s in enumerate(subs):
            if isinstance(s, tuple) and s[0] == "slice":
                lo = self.eval(s[1], frame) if s[1] is not None else arr.lb[d]
                hi = self.eval(s[2], frame) if s[2] is not None else arr.lb[d] + arr.a.shape[d] - 1
                st = self.eval(s[3], frame) if s[3] is not None else 1
                if lo < arr.lb[d] or hi > arr.lb[d] + arr.a.shape[d] - 1:
                    if hi >= lo:
                        raise IndexError(f"section {lo}:{hi} out of bounds {arr.lb[d]}:{arr.lb[d] + arr.a.shape[d] - 1}")
                idx.append(slice(lo - arr.lb[d], hi - arr.lb[d] + 1, st))
                is_sec = True
            else:
                v = self.eval(s, frame)
                if isinstance(v, np.ndarray):
                    idx.append(v - arr.lb[d])
                    is_sec = True
                else:
                    v = int(v)
                    if v < arr.lb[d] or v > arr.lb[d] + arr.a.shape[d] - 1:
                        raise IndexError(f"index {v} out of bounds {arr.lb[d]}:{arr.lb[d] + arr.a.shape[d] - 1} (dim {d + 1})")
                    idx.append(v - arr.lb[d])
        return tuple(idx), is_sec

    def assign(self, lhs, val, frame):
        if isinstance(val, FArray):
            val = val.a
        if lhs[0] == "var":
            cur = None
            try:
                cur = self.lookup(lhs[1], frame)
            except NameError:
                pass
            if isinstance(cur, FArray):
                cur.a[...] = val
            else:
                if isinstance(cur, float) and isinstance(val, (int, bool)) and not isinstance(val, bool):
                    val = float(val)
                if isinstance(cur, int) and not isinstance(cur, bool) and isinstance(val, float):
                    val = int(val)
                self.store(lhs[1], val, frame)
        elif lhs[0] == "call":
            arr = self.lookup(lhs[1], frame)
            idx, _ = self.index_of(arr, lhs[2], frame)
            arr.a[idx] = val
        else:
            raise NotImplementedError(f"assignment target {lhs}")

    def eval(self, n, frame):
        k = n[0]
        if k == "num":
            return n[1]
        if k == "var":
            v = self.lookup(n[1], frame)
            if v is None:
                raise NameError(f"use of undefined {n[1]!r} in {frame['proc'].name}")
            return v.a if isinstance(v, FArray) else v
        if k == "paren":
            return self.eval(n[1], frame)
        if k == "neg":
            return -self.eval(n[1], frame)
        if k == "not":
            return not self.truth(self.eval(n[1], frame))
        if k == "bin":
            op = n[1]
            if op == ".and.":
                return self.truth(self.eval(n[2], frame)) and self.truth(self.eval(n[3], frame))
            if op == ".or.":
                return self.truth(self.eval(n[2], frame)) or self.truth(self.eval(n[3], frame))
            a, b = self.eval(n[2], frame), self.eval(n[3], frame)
            if op == "+":
                return a + b
            if op == "-":
                return a - b
            if op == "*":
                return a * b
            if op == "/":
                if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)) and not isinstance(a, bool):
                    q = abs(a) // abs(b)
                    return int(q if (a >= 0) == (b >= 0) else -q)
                return a / b
            if op == "**":
                return self.power(a, b)
            if op == "==":
                return a == b
            if op == "/=":
                return a != b
            if op == "<":
                return a < b
            if op == "<=":
                return a <= b
            if op == ">":
                return a > b
            if op == ">=":
                return a >= b
            if op == ".eqv.":
                return bool(a) == bool(b)
            if op == ".neqv.":
                return bool(a) != bool(b)
            raise NotImplementedError(op)
        if k == "call":
            name = n[1]
            try:
                v = self.lookup(name, frame)
            except NameError:
                v = None
            if isinstance(v, FArray):
                idx, is_sec = self.index_of(v, n[2], frame)
                r = v.a[idx]
                if is_sec:
                    return r
                return r.item() if isinstance(r, np.generic) else r
            if name in self.procs and self.procs[name].kind == "function":
                return self._invoke(name, [self.eval_arg(a, frame) for a in n[2]], {}, None)
            if name in self.intr:
                return self.intr[name](*[self.eval(a, frame) for a in n[2]])
            if name in ("size",):
                arr = self.lookup(n[2][0][1], frame)
                if len(n[2]) > 1:
                    return int(arr.a.shape[int(self.eval(n[2][1], frame)) - 1])
                return int(arr.a.size)
            if name == "allocated":
                return self.lookup(n[2][0][1], frame) is not None
            if name == "present":      # optional dummy: passed iff bound to a value in the callee's frame
                return frame["loc"].get(n[2][0][1]) is not None
            raise NameError(f"unknown function/array {name!r} in {frame['proc'].name}")
        if k == "arrcon":
            return np.array([self.eval(x, frame) for x in n[1]])
        if k == "str":
            return n[1]
        raise NotImplementedError(n)

    @staticmethod
    def power(a, b):
        # gfortran: integer (or integral real) exponents -> repeated multiplication
        if isinstance(b, float) and b == int(b) and abs(b) <= 8:
            b = int(b)
        if isinstance(b, (int, np.integer)) and not isinstance(b, bool):
            if isinstance(a, (int, np.integer)) and b >= 0:
                return int(a) ** int(b)
            r = a
            for _ in range(abs(int(b)) - 1):
                r = r * a
            if b == 0:
                return 1.0
            return r if b > 0 else 1.0 / r
        if isinstance(a, np.ndarray):
            return np.power(a, b)
        return math.pow(a, b)
