"""ORACLE (test infrastructure only -- never imported by the product path).

Parser for the subset of 2013-era Julia surface syntax that SimpleMCMC.jl models and
derivation rules are written in.  It produces the same tree the Julia parser handed to
`parseModel!` (/root/reference/src/parsing.jl:76-120): `Expr(head, args)` nodes with heads
`block = += -= *= call ref ' vcat comparison :` and `Sym` leaves, with n-ary `+`/`*` calls
(flattened the way Julia's reader does, later re-nested by `unfold!`, parsing.jl:160-166).
"""
import re


class Sym(str):
    """A Julia Symbol."""
    __slots__ = ()

    def __repr__(self):
        return ":" + str(self)


class Expr(object):
    __slots__ = ("head", "args", "_nary")

    def __init__(self, head, *args):
        self.head = head
        self.args = list(args)
        self._nary = False   # parser-internal: an n-ary +/* chain still open for merging

    def __repr__(self):
        return show(self)

    def __eq__(self, o):
        return isinstance(o, Expr) and self.head == o.head and self.args == o.args

    def __hash__(self):
        return hash((self.head, len(self.args)))

    def copy(self):
        return Expr(self.head, *[a.copy() if isinstance(a, Expr) else a for a in self.args])


_TOK = re.compile(r"""
    (?P<ws>[ \t]+)
  | (?P<comment>\#[^\n]*)
  | (?P<nl>[\n;])
  | (?P<num>(?:\d+\.\d*(?![*/^\\])|\d+\.(?=[*/^\\])|\.\d+|\d+)(?:[eE][-+]?\d+)?)
  | (?P<id>[A-Za-z_][A-Za-z_0-9!]*)
  | (?P<op>\.<=|\.>=|\.==|\.\*|\./|\.\^|\.<|\.>|\+=|-=|\*=|<=|>=|==|[-+*/^<>=~'(),\[\]:])
""", re.X)


def tokenize(src):
    out, pos = [], 0
    while pos < len(src):
        m = _TOK.match(src, pos)
        if not m:
            raise SyntaxError("oracle parser: bad character %r at %d" % (src[pos], pos))
        kind = m.lastgroup
        txt = m.group(kind)
        if kind == "num":
            # "1./x" is `1 ./ x` (the trailing dot belongs to the operator)
            if txt.endswith(".") and src[m.end():m.end() + 1] in ("/", "*", "^"):
                txt = txt[:-1]
                pos = m.end() - 1
            else:
                pos = m.end()
            isint = re.fullmatch(r"\d+", txt) is not None
            out.append(("num", int(txt) if isint else float(txt), m.start() > 0 and src[m.start() - 1] in " \t"))
            continue
        if kind not in ("ws", "comment"):
            out.append((kind, txt, m.start() > 0 and src[m.start() - 1] in " \t"))
        pos = m.end()
    out.append(("eof", None, False))
    return out


_CMP = {"<", ">", "<=", ">=", "==", ".<", ".>", ".<=", ".>=", ".=="}


class _Parser(object):
    def __init__(self, src):
        self.t = tokenize(src)
        self.i = 0
        self.in_ref = 0

    def peek(self):
        return self.t[self.i]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def accept(self, kind, val=None):
        k, v, _ = self.peek()
        if k == kind and (val is None or v == val):
            self.i += 1
            return True
        return False

    def expect(self, kind, val):
        if not self.accept(kind, val):
            raise SyntaxError("oracle parser: expected %r, found %r" % (val, self.peek()[:2]))

    # block := stmt (sep stmt)*
    def block(self, closers=("eof",)):
        stmts = []
        while True:
            while self.accept("nl"):
                pass
            k, v, _ = self.peek()
            if k == "eof" or (k == "op" and v in closers):
                break
            stmts.append(self.statement())
        return Expr("block", *stmts)

    def statement(self):
        lhs = self.tilde()
        k, v, _ = self.peek()
        if k == "op" and v in ("=", "+=", "-=", "*="):
            self.next()
            rhs = self.statement()
            return Expr(v, lhs, rhs)
        return lhs

    def tilde(self):
        lhs = self.comparison()
        if self.accept("op", "~"):
            rhs = self.comparison()
            return Expr("call", Sym("~"), lhs, rhs)
        return lhs

    def comparison(self):
        first = self.range_()
        parts = [first]
        while self.peek()[0] == "op" and self.peek()[1] in _CMP:
            parts.append(Sym(self.next()[1]))
            parts.append(self.range_())
        return first if len(parts) == 1 else Expr("comparison", *parts)

    def range_(self):
        k, v, _ = self.peek()
        if self.in_ref and k == "op" and v == ":":
            # a bare ':' inside a ref
            nk, nv, _ = self.t[self.i + 1]
            if nk == "op" and nv in (",", "]"):
                self.next()
                return Sym(":")
        a = self.additive()
        if self.peek()[:2] == ("op", ":"):
            self.next()
            b = self.additive()
            return Expr(":", a, b)
        return a

    def additive(self):
        node = self.multiplicative()
        while self.peek()[0] == "op" and self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            rhs = self.multiplicative()
            if op == "+" and isinstance(node, Expr) and node.head == "call" and node.args[0] == "+" \
                    and node._nary:
                node.args.append(rhs)
            else:
                node = Expr("call", Sym(op), node, rhs)
                if op == "+":
                    node._nary = True
        return node

    def multiplicative(self):
        node = self.unary()
        while self.peek()[0] == "op" and self.peek()[1] in ("*", "/", ".*", "./"):
            op = self.next()[1]
            rhs = self.unary()
            if op == "*" and isinstance(node, Expr) and node.head == "call" and node.args[0] == "*" \
                    and node._nary:
                node.args.append(rhs)
            else:
                node = Expr("call", Sym(op), node, rhs)
                if op == "*":
                    node._nary = True
        return node

    def unary(self):
        k, v, _ = self.peek()
        if k == "op" and v in ("-", "+"):
            self.next()
            nk, nv, _ = self.peek()
            if nk == "num":
                # Julia folds a sign into an adjacent numeric literal unless `^` follows it
                save = self.i
                self.next()
                k2, v2, _ = self.peek()
                self.i = save
                if not (k2 == "op" and v2 in ("^", ".^")):
                    lit = self.power()   # handles juxtaposition such as -2x too
                    if isinstance(lit, (int, float)) and not isinstance(lit, bool):
                        return -lit if v == "-" else lit
                    return Expr("call", Sym(v), lit)
            operand = self.unary()
            return Expr("call", Sym(v), operand)
        return self.power()

    def power(self):
        base = self.postfix()
        k, v, _ = self.peek()
        if k == "op" and v in ("^", ".^"):
            self.next()
            expo = self.unary()  # right associative, allows 2^-x
            return Expr("call", Sym(v), base, expo)
        return base

    def postfix(self):
        node = self.atom()
        while True:
            k, v, spaced = self.peek()
            if k == "op" and v == "(" and not spaced and isinstance(node, Sym):
                self.next()
                args = self.arglist(")")
                node = Expr("call", node, *args)
            elif k == "op" and v == "[" and not spaced:
                self.next()
                self.in_ref += 1
                args = self.arglist("]")
                self.in_ref -= 1
                node = Expr("ref", node, *args)
            elif k == "op" and v == "'" and not spaced:
                self.next()
                node = Expr("'", node)
            else:
                break
        return node

    def arglist(self, closer):
        args = []
        saved, self.in_ref = self.in_ref, (self.in_ref if closer == "]" else 0)
        if not self.accept("op", closer):
            while True:
                args.append(self.tilde())
                if self.accept("op", closer):
                    break
                self.expect("op", ",")
        self.in_ref = saved
        return args

    def atom(self):
        k, v, _ = self.next()
        if k == "num":
            # numeric-literal coefficient: 2(mu-x), 2lmu
            nk, nv, spaced = self.peek()
            if not spaced and ((nk == "op" and nv == "(") or nk == "id"):
                rhs = self.postfix() if nk == "id" else self.atom_paren()
                node = Expr("call", Sym("*"), v, rhs)
                return node
            return v
        if k == "id":
            if v == "true":
                return True
            if v == "false":
                return False
            return Sym(v)
        if k == "op" and v == "(":
            self.i -= 1
            return self.atom_paren()
        if k == "op" and v == "[":
            # vcat literal [a; b] / [a, b]
            items = []
            while not self.accept("op", "]"):
                items.append(self.tilde())
                if not (self.accept("op", ",") or self.accept("nl")):
                    self.expect("op", "]")
                    break
            return Expr("vcat", *items)
        raise SyntaxError("oracle parser: unexpected token %r" % ((k, v),))

    def atom_paren(self):
        self.expect("op", "(")
        saved, self.in_ref = self.in_ref, 0
        blk = self.block(closers=(")",))
        self.expect("op", ")")
        self.in_ref = saved
        if len(blk.args) == 1:
            inner = blk.args[0]
            if isinstance(inner, Expr):
                inner._nary = False   # parenthesised sums are not merged into an outer n-ary call
            return inner
        return blk


def _fix(node):
    if isinstance(node, Expr):
        for a in node.args:
            _fix(a)
    return node


def parse(src):
    """Parse a model (the inside of `quote ... end`) into Expr(:block, ...)."""
    p = _Parser(src)
    blk = p.block()
    return _fix(blk)


def parse_expr(src):
    blk = parse(src)
    assert len(blk.args) == 1, "expected a single expression"
    return blk.args[0]


def show(e):
    if isinstance(e, Expr):
        h = e.head
        if h == "call":
            f = e.args[0]
            if str(f) in ("+", "-", "*", "/", ".*", "./", "^", ".^", "~") and len(e.args) >= 3:
                return "(" + (" %s " % f).join(show(a) for a in e.args[1:]) + ")"
            return "%s(%s)" % (f, ", ".join(show(a) for a in e.args[1:]))
        if h in ("=", "+=", "-=", "*="):
            return "%s %s %s" % (show(e.args[0]), h, show(e.args[1]))
        if h == "ref":
            return "%s[%s]" % (show(e.args[0]), ",".join(show(a) for a in e.args[1:]))
        if h == ":":
            return "%s:%s" % (show(e.args[0]), show(e.args[1]))
        if h == "'":
            return show(e.args[0]) + "'"
        if h == "comparison":
            return "(" + " ".join(show(a) for a in e.args) + ")"
        if h == "block":
            return "begin " + "; ".join(show(a) for a in e.args) + " end"
        if h == "vcat":
            return "[" + "; ".join(show(a) for a in e.args) + "]"
        return "Expr(%s, %s)" % (h, ", ".join(show(a) for a in e.args))
    if isinstance(e, Sym):
        return str(e)
    return repr(e)
