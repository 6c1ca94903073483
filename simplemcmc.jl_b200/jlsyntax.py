"""Reader for the model DSL: the Julia surface syntax accepted inside `quote ... end` by SimpleMCMC.jl
(/root/reference/README.md:56-104; /root/reference/src/parsing.jl:76-120 lists the statement forms).

Produces small tuples:  ("num", v) ("var", name) ("call", fname, [args]) ("ref", base, [indices])
("range", a, b) ("colon",) ("end",) ("assign", op, lhs, rhs) ("tilde", lhs, distname, [args]) ("cmp", [a, op, b, ...])
("block", [stmts]).  Binary + and * chains are n-ary, as in Julia's reader.
"""
import re

_TOKEN = re.compile(
    r"\s*(?:(?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?)|(?P<name>[A-Za-z_]\w*)|"
    r"(?P<op>\.<=|\.>=|\.\*|\./|\.\^|\.<|\.>|\+=|-=|\*=|<=|>=|==|[-+*/^<>=~'(),\[\]:;\n]))", re.S)

_COMPARE = ("<", ">", "<=", ">=", "==", ".<", ".>", ".<=", ".>=")


class DSLSyntaxError(Exception):
    pass


def _lex(text):
    text = re.sub(r"#[^\n]*", "", text)
    toks, pos = [], 0
    text = text.replace("\r", "")
    while True:
        # keep newlines as separators but skip other blanks
        m = re.compile(r"[ \t]*").match(text, pos)
        pos = m.end()
        if pos >= len(text):
            break
        if text[pos] == "\n":
            toks.append(("op", "\n", True))
            pos += 1
            continue
        m = _TOKEN.match(text, pos)
        if not m or m.end() == pos:
            raise DSLSyntaxError("cannot read model text at: %r" % text[pos:pos + 20])
        glued = pos > 0 and text[pos - 1] not in " \t\n"
        if m.group("num") is not None:
            s = m.group("num")
            end = m.end()
            if s.endswith(".") and text[end:end + 1] in ("*", "/", "^"):   # `1./x` reads as 1 ./ x
                s, end = s[:-1], end - 1
            toks.append(("num", float(s) if re.search(r"[.eE]", s) else int(s), glued))
            pos = end
            continue
        kind = "name" if m.group("name") is not None else "op"
        toks.append((kind, m.group(kind), glued))
        pos = m.end()
    toks.append(("eof", None, False))
    return toks


class _Reader(object):
    def __init__(self, text):
        self.toks = _lex(text)
        self.k = 0
        self.bracket = 0

    def at(self, kind, val=None):
        t = self.toks[self.k]
        return t[0] == kind and (val is None or t[1] == val)

    def at_op(self, *vals):
        t = self.toks[self.k]
        return t[0] == "op" and t[1] in vals

    def take(self):
        t = self.toks[self.k]
        self.k += 1
        return t

    def need(self, val):
        if not self.at_op(val):
            raise DSLSyntaxError("expected %r near token %r" % (val, self.toks[self.k][1]))
        self.k += 1

    def statements(self, stop=None):
        out = []
        while True:
            while self.at_op("\n", ";"):
                self.k += 1
            if self.at("eof") or (stop and self.at_op(stop)):
                return out
            out.append(self.statement())

    def statement(self):
        lhs = self.expr()
        if self.at_op("=", "+=", "-=", "*="):
            op = self.take()[1]
            rhs = self.statement()
            return ("assign", op, lhs, rhs)
        return lhs

    def expr(self):
        lhs = self.compare()
        if self.at_op("~"):
            self.k += 1
            rhs = self.compare()
            if rhs[0] != "call":
                raise DSLSyntaxError("the right-hand side of ~ must be Distribution(args...)")
            return ("tilde", lhs, rhs[1], rhs[2])
        return lhs

    def compare(self):
        parts = [self.span()]
        while self.at_op(*_COMPARE):
            parts.append(self.take()[1])
            parts.append(self.span())
        return parts[0] if len(parts) == 1 else ("cmp", parts)

    def span(self):
        if self.bracket and self.at_op(":"):
            nxt = self.toks[self.k + 1]
            if nxt[0] == "op" and nxt[1] in (",", "]"):
                self.k += 1
                return ("colon",)
        a = self.sum_()
        if self.at_op(":"):
            self.k += 1
            return ("range", a, self.sum_())
        return a

    def sum_(self):
        node, open_nary = self.product(), False
        while self.at_op("+", "-"):
            op = self.take()[1]
            rhs = self.product()
            if op == "+" and open_nary:
                node[2].append(rhs)
            else:
                node = ("call", op, [node, rhs])
                open_nary = (op == "+")
        return node

    def product(self):
        node, open_nary = self.prefix(), False
        while self.at_op("*", "/", ".*", "./"):
            op = self.take()[1]
            rhs = self.prefix()
            if op == "*" and open_nary:
                node[2].append(rhs)
            else:
                node = ("call", op, [node, rhs])
                open_nary = (op == "*")
        return node

    def prefix(self):
        if self.at_op("-", "+"):
            op = self.take()[1]
            if self.at("num"):
                nxt = self.toks[self.k + 1]
                if not (nxt[0] == "op" and nxt[1] in ("^", ".^")):
                    v = self.power()
                    if v[0] == "num":
                        return ("num", -v[1] if op == "-" else v[1])
                    return ("call", op, [v])
            return ("call", op, [self.prefix()])
        return self.power()

    def power(self):
        base = self.postfix()
        if self.at_op("^", ".^"):
            op = self.take()[1]
            return ("call", op, [base, self.prefix()])
        return base

    def postfix(self):
        node = self.primary()
        while True:
            t = self.toks[self.k]
            if t[0] == "op" and t[1] == "(" and t[2] and node[0] == "var":
                self.k += 1
                node = ("call", node[1], self.args(")"))
            elif t[0] == "op" and t[1] == "[" and t[2]:
                self.k += 1
                self.bracket += 1
                idx = self.args("]")
                self.bracket -= 1
                node = ("ref", node, idx)
            elif t[0] == "op" and t[1] == "'" and t[2]:
                self.k += 1
                node = ("call", "transpose", [node])
            else:
                return node

    def args(self, closer):
        out = []
        saved = self.bracket
        if closer == ")":
            self.bracket = 0
        if self.at_op(closer):
            self.k += 1
        else:
            while True:
                out.append(self.expr())
                if self.at_op(closer):
                    self.k += 1
                    break
                self.need(",")
        self.bracket = saved
        return out

    def primary(self):
        kind, val, _ = self.take()
        if kind == "num":
            t = self.toks[self.k]
            if t[2] and (t[0] == "name" or (t[0] == "op" and t[1] == "(")):   # 2x, 2(a+b)
                return ("call", "*", [("num", val), self.postfix()])
            return ("num", val)
        if kind == "name":
            if val == "end" and self.bracket:
                return ("end",)
            if val in ("true", "false"):
                return ("num", 1 if val == "true" else 0)
            return ("var", val)
        if kind == "op" and val == "(":
            saved, self.bracket = self.bracket, 0
            body = self.statements(stop=")")
            self.need(")")
            self.bracket = saved
            if len(body) == 1:
                inner = body[0]
                if inner[0] == "call" and inner[1] in ("+", "*"):
                    inner = ("call", inner[1], list(inner[2]), "closed")
                return inner
            return ("block", body)
        if kind == "op" and val == "[":
            raise DSLSyntaxError("array literals / vcat are parsed by the reference but have no derivation rule "
                                 "(src/parsing.jl:130, src/diff.jl:7)")
        raise DSLSyntaxError("unexpected %r in model text" % (val,))


def read_model(text):
    """model text -> list of statements"""
    r = _Reader(text)
    return _strip(r.statements())


def _strip(node):
    # drop the "closed" marker used to stop n-ary merging across parentheses
    if isinstance(node, list):
        return [_strip(n) for n in node]
    if isinstance(node, tuple):
        if node and node[0] == "call":
            return ("call", node[1], [_strip(a) for a in node[2]])
        return tuple(_strip(n) if isinstance(n, (tuple, list)) else n for n in node)
    return node
