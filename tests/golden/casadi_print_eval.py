"""Evaluator for the expression strings CasADi prints into acados' OCP dump (test infrastructure).

`acados_ocp_SNMPC.json` of the reference holds the folded SNMPC model as printed MX expressions:
`model.disc_dyn_expr` (json line 775), `model.con_h_expr` (:745), `model.cost_y_expr` (:772) and their
terminal twins.  This module parses that print syntax into an AST and evaluates it with numpy, so that the
SNMPC oracle can be pinned against the reference's OWN model rather than against itself:

    @k=<expr>, @k+1=<expr>, ..., <final expr>          definitions, then the value
    (c?a:b)   !c   a<b   a==b   + - * /   -a           scalar / element-wise operators
    sq sqrt sin cos atan asin fmax fmin fmod pow       element-wise functions
    vertcat(...)  mac(X,Y,Z)=X@Y+Z  reshape(v)  X'  X[i]  X[a:b]  X[a:b:s]  zeros(RxC[,Nnz])  ones(RxC)
    dense(X) project(X)                                sparsity casts: identity here
    (X[i] = e)                                         copy of X with entry i replaced
    f(x, u){0}                                         call of the CasADi Function `f` built from the RHS
    ax_max_interpolant(v){0}, ay_max_interpolant(v){0} the g-g-v look-ups (casadi.interpolant 'linear')

Nothing is imported from the reference: the parser reads the JSON text only.  Matrices are 2-D numpy arrays
(column vectors are (n, 1)); linear indices are column-major like CasADi's.  The print drops the target shape
of `reshape`; it is recovered from the element count and the reference source that built the expression
(pred_model_dynamic_disc.py:213: reshape(f_disc, 8, n_samples); SNMPC_acados_settings.py:112 and
pred_model_dynamic_disc.py:61: reshape(p[:-2], n_samples, num_poly_terms)).

Numeric literals are printed with 6 significant digits.  `Evaluator(..., literals=...)` lets the caller replace
a printed literal by the full-precision constant it is the rounding of (after asserting that it IS its
rounding), which turns a 1e-5 comparison into a 1e-12 one.
"""
import math
import re

import numpy as np

_TOK = re.compile(r"\s*(?:(@\d+)|(\d+x\d+)|(\d+nz)|(\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?)|([A-Za-z_]\w*)|(==|[()\[\]{},?:!<=+\-*/']))")


def tokenize(s):
    out, pos = [], 0
    while pos < len(s):
        m = _TOK.match(s, pos)
        if not m:
            if s[pos:].strip() == "":
                break
            raise SyntaxError(f"cannot tokenize at {pos}: {s[pos:pos + 40]!r}")
        pos = m.end()
        ref, dims, nnz, num, name, sym = m.groups()
        if ref: out.append(("ref", int(ref[1:])))
        elif dims: out.append(("dims", tuple(int(v) for v in dims.split("x"))))
        elif nnz: out.append(("nnz", int(nnz[:-2])))
        elif num: out.append(("num", num))
        elif name: out.append(("name", name))
        else: out.append(("sym", sym))
    out.append(("end", None))
    return out


class Parser:
    """Pratt parser.  Precedence (low -> high): '=' (entry assignment), '?:', '==', '<', '+ -', '* /', unary, postfix."""

    def __init__(self, text):
        self.t = tokenize(text)
        self.i = 0

    def peek(self, k=0):
        return self.t[self.i + k]

    def take(self, kind=None, val=None):
        tok = self.t[self.i]
        if (kind and tok[0] != kind) or (val is not None and tok[1] != val):
            raise SyntaxError(f"expected {kind} {val}, got {tok} at token {self.i}")
        self.i += 1
        return tok

    def at(self, val):
        return self.t[self.i] == ("sym", val)

    def program(self):
        defs = {}
        while self.peek()[0] == "ref" and self.peek(1) == ("sym", "="):
            k = self.take("ref")[1]
            self.take("sym", "=")
            defs[k] = self.expr(1)
            self.take("sym", ",")
        final = self.expr(0)
        self.take("end")
        return defs, final

    BIN = {"==": 3, "<": 4, "+": 5, "-": 5, "*": 6, "/": 6}

    def expr(self, minp):
        lhs = self.unary()
        while True:
            tok = self.peek()
            if tok[0] != "sym":
                break
            op = tok[1]
            if op == "=" and minp <= 0:
                self.take()
                lhs = ("assign", lhs, self.expr(1))
            elif op == "?" and minp <= 2:
                self.take()
                a = self.expr(1)
                self.take("sym", ":")
                b = self.expr(2)
                lhs = ("ifelse", lhs, a, b)
            elif op in self.BIN and self.BIN[op] >= max(minp, 3):
                self.take()
                lhs = ("bin", op, lhs, self.expr(self.BIN[op] + 1))
            else:
                break
        return lhs

    def unary(self):
        if self.at("-"):
            self.take()
            return ("neg", self.unary())
        if self.at("!"):
            self.take()
            return ("not", self.unary())
        return self.postfix(self.primary())

    def postfix(self, node):
        while True:
            if self.at("'"):
                self.take()
                node = ("T", node)
            elif self.at("{"):
                self.take(); self.take("num"); self.take("sym", "}")      # output 0 of a Function call
            elif self.at("["):
                self.take()
                parts, cur = [], None
                while not self.at("]"):
                    if self.at(":"):
                        self.take(); parts.append(cur); cur = None
                    else:
                        cur = self.expr(1)
                parts.append(cur)
                self.take("sym", "]")
                node = ("index", node, parts)
            else:
                return node

    def primary(self):
        tok = self.take()
        if tok[0] == "num":
            return ("num", tok[1])
        if tok[0] == "ref":
            return ("ref", tok[1])
        if tok == ("sym", "("):
            e = self.expr(0)
            self.take("sym", ")")
            return e
        if tok[0] == "name":
            if self.at("("):
                self.take()
                args = []
                while not self.at(")"):
                    if self.peek()[0] in ("dims", "nnz"):
                        args.append(self.take())
                    else:
                        args.append(self.expr(1))
                    if self.at(","):
                        self.take()
                self.take("sym", ")")
                return ("call", tok[1], args)
            return ("var", tok[1])
        raise SyntaxError(f"unexpected token {tok} at {self.i}")


def _lin(X):
    return np.asarray(X, float).reshape(-1, order="F")


def interp_linear(xs, ys, x):
    """casadi.interpolant('linear'): piecewise linear, linearly extrapolated from the end intervals."""
    xs, ys = np.asarray(xs, float), np.asarray(ys, float)
    i = int(np.clip(np.searchsorted(xs, x, side="right") - 1, 0, len(xs) - 2))
    return ys[i] + (ys[i + 1] - ys[i]) / (xs[i + 1] - xs[i]) * (x - xs[i])


class Evaluator:
    """Evaluates a parsed program.  `env`: variable name -> value.  `functions`: name -> python callable.
    `literals`: printed literal text -> replacement value (see module docstring)."""

    ELEM = {"sq": lambda a: a * a, "sqrt": np.sqrt, "sin": np.sin, "cos": np.cos, "atan": np.arctan, "asin": np.arcsin}

    def __init__(self, defs, env, functions=None, literals=None, reshape_rows=None):
        self.defs, self.env, self.fn = defs, env, functions or {}
        self.lit = literals or {}
        self.reshape_rows = reshape_rows or {}
        self.cache = {}
        self.used_literals = set()

    def child(self, env):
        e = Evaluator(self.defs, env, self.fn, self.lit, self.reshape_rows)
        e.used_literals = self.used_literals
        return e

    def ev(self, n):
        k = n[0]
        if k == "num":
            self.used_literals.add(n[1])
            return self.lit[n[1]] if n[1] in self.lit else float(n[1])
        if k == "ref":
            if n[1] not in self.cache:
                self.cache[n[1]] = self.ev(self.defs[n[1]])
            return self.cache[n[1]]
        if k == "var":
            return self.env[n[1]]
        if k == "neg":
            return -self.ev(n[1])
        if k == "not":
            return 0.0 if self.ev(n[1]) != 0 else 1.0
        if k == "T":
            return np.asarray(self.ev(n[1])).T
        if k == "bin":
            a, b = self.ev(n[2]), self.ev(n[3])
            op = n[1]
            if op == "+": return a + b
            if op == "-": return a - b
            if op == "*": return a * b
            if op == "/": return a / b
            if op == "<": return 1.0 if a < b else 0.0
            if op == "==": return 1.0 if a == b else 0.0
        if k == "ifelse":      # CasADi's if_else evaluates both branches; so do we (no short circuit)
            c, a, b = self.ev(n[1]), self.ev(n[2]), self.ev(n[3])
            return a if c != 0 else b
        if k == "index":
            X, parts = self.ev(n[1]), [None if p is None else int(self.ev(p)) for p in n[2]]
            v = _lin(X)
            if len(parts) == 1:
                return float(v[parts[0]])
            return v[slice(*parts)].reshape(-1, 1)
        if k == "assign":
            tgt = n[1]
            assert tgt[0] == "index" and len(tgt[2]) == 1
            X = np.array(self.ev(tgt[1]), float)
            shape = X.shape
            v = _lin(X).copy()
            v[int(self.ev(tgt[2][0]))] = self.ev(n[2])
            return v.reshape(shape, order="F")
        if k == "call":
            return self.call(n[1], n[2])
        raise ValueError(k)

    def call(self, name, args):
        if name in ("zeros", "ones"):
            r, c = args[0][1]
            return np.zeros((r, c)) if name == "zeros" else np.ones((r, c))
        vals = [self.ev(a) for a in args]
        if name in self.ELEM:
            return self.ELEM[name](vals[0])
        if name == "fmax": return np.maximum(vals[0], vals[1])
        if name == "fmin": return np.minimum(vals[0], vals[1])
        if name == "fmod": return math.fmod(vals[0], vals[1])
        if name == "pow": return vals[0] ** vals[1]
        if name in ("dense", "project"): return vals[0]
        if name == "vertcat":
            # (the print of a row slice `X[0, :]` is a linear slice: vectors of either orientation stack as columns)
            return np.vstack([np.asarray(v, float).reshape(-1, 1) if (np.ndim(v) < 2 or 1 in np.shape(v)) else v for v in vals])
        if name == "mac":
            return np.asarray(vals[0]) @ np.asarray(vals[1]) + vals[2]
        if name == "reshape":
            v = _lin(vals[0])
            rows = self.reshape_rows[v.size]
            return v.reshape((rows, v.size // rows), order="F")
        if name in self.fn:
            return self.fn[name](self, *vals)
        raise NameError(f"unknown function {name}")


def parse(text):
    return Parser(text).program()
