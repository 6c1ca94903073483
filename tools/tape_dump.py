"""Prints a device op tape (include/simplemcmc_tape.h) in readable form: header, registers, statements, bound data, index
tables, register programs.  For whoever writes another emitter (the Julia side of INTEGRATION.md): dump your blob and
`tests/golden/<model>.tape`, and diff the text.
    python tools/tape_dump.py tests/golden/sleepstudy.tape
`parse(blob)` returns the same content as a dict (tests/test_golden_tapes.py walks the golden blobs through it)."""
import struct
import sys

KIND = ["imm", "data", "hdr", "chain", "param", "grad"]
OPS = {0: "COPY", 1: "NEG", 2: "EXP", 3: "LOG", 4: "SIN", 5: "COS", 6: "ABS", 7: "SIGN", 8: "DIGAMMA", 9: "ZERO", 10: "SQRT",
       16: "ADD", 17: "SUB", 18: "MUL", 19: "DIV", 20: "POW", 21: "MAX", 22: "MIN", 23: "GT", 24: "LT", 25: "GE", 26: "LE", 27: "EQ",
       32: "LOGPDF", 33: "DLOGPDF", 40: "MATMUL", 41: "TRANSPOSE", 42: "GATHER", 43: "SCATTER_SET", 44: "SCATTER_ADD", 48: "GLM",
       49: "FUSED", 50: "AFFNORM", 100: "loadv", 101: "loads", 102: "storev", 103: "reduce"}
DISTS = ["Normal", "Uniform", "Weibull", "Beta", "TDist", "Exponential", "Gamma", "Cauchy", "logNormal", "Bernoulli", "Binomial",
         "Poisson", "TestDiff"]
GLM = ["NORMAL", "BERNOULLI", "BINOMIAL", "POISSON"]
FLAGS = [(1, "+="), (2, "tA"), (4, "tB"), (8, "fwd-only")]


def parse(blob):
    """the blob, exactly as the header lays it out; raises ValueError when bytes are missing or left over"""
    try:
        return _parse(blob)
    except struct.error as e:
        raise ValueError("truncated tape: %s" % e)


def _parse(blob):
    if len(blob) < 48:
        raise ValueError("shorter than a tape header")
    magic, version, n_regs, n_stmts, n_params, n_data, n_idx, result, flags, n_fwd, n_prog, _ = struct.unpack("<12I", blob[:48])
    if magic != 0x54434D53:
        raise ValueError("bad magic %08x" % magic)
    t = dict(version=version, n_params=n_params, result_reg=result, flags=flags, n_fwd=n_fwd, regs=[], stmts=[], data=[], tables=[],
             programs=[])
    p = 48
    for _ in range(n_regs):
        kind, rows, cols, ndim, aux, imm = struct.unpack("<4Iqd", blob[p:p + 32])
        t["regs"].append(dict(kind=kind, rows=rows, cols=cols, ndim=ndim, aux=aux, imm=imm))
        p += 32
    for _ in range(n_stmts):
        op, fl, dst, nsrc, s0, s1, s2, s3, a0, a1, a2, _pad, imm = struct.unpack("<4I4I4Id", blob[p:p + 56])
        t["stmts"].append(dict(op=op, flags=fl, dst=dst, src=[s0, s1, s2, s3][:nsrc], aux0=a0, aux1=a1, aux2=a2, imm=imm))
        p += 56
    for _ in range(n_data):
        nm, rows, cols = struct.unpack("<56s2I", blob[p:p + 64])
        t["data"].append(dict(name=nm.rstrip(b"\0").decode(), rows=rows, cols=cols))
        p += 64
    for _ in range(n_idx):
        n, _pad = struct.unpack("<2I", blob[p:p + 8])
        t["tables"].append(list(struct.unpack("<%di" % n, blob[p + 8:p + 8 + 4 * n])))
        p += 8 + 4 * ((n + 1) & ~1)
    for _ in range(n_prog):
        n_instr, n_pro, n_red, n = struct.unpack("<4I", blob[p:p + 16])
        p += 16
        instr = []
        for _i in range(n_instr):
            op, dst, nsrc, s0, s1, s2, fl, a0, a1, reg = struct.unpack("<HBB3BBHHI", blob[p:p + 16])
            instr.append(dict(op=op, dst=dst, src=[s0, s1, s2][:nsrc], flags=fl, aux0=a0, aux1=a1, reg=reg))
            p += 16
        red = []
        for _i in range(n_red):
            red.append(struct.unpack("<2I", blob[p:p + 8]))
            p += 8
        t["programs"].append(dict(n=n, n_prologue=n_pro, instr=instr, red=red))
    if p != len(blob):
        raise ValueError("%d bytes left over" % (len(blob) - p))
    return t


def _reg(t, r):
    g = t["regs"][r]
    k = KIND[g["kind"]]
    if k == "imm":
        return "%r" % g["imm"]
    shape = "" if g["ndim"] == 0 else ("[%d]" % g["rows"] if g["ndim"] == 1 else "[%d,%d]" % (g["rows"], g["cols"]))
    if k == "data":
        return "r%d:%s%s" % (r, t["data"][g["aux"]]["name"], shape)
    if k in ("param", "grad"):
        return "r%d:%s@%d%s" % (r, "beta" if k == "param" else "grad", g["aux"], shape)
    return "r%d:%s%s" % (r, k, shape)


def _flags(fl):
    return "".join(" " + name for bit, name in FLAGS if fl & bit and name != "+=")


def lines(t):
    out = ["tape version %d: %d parameters, %d registers, %d statements (%d forward), result r%d" % (
        t["version"], t["n_params"], len(t["regs"]), len(t["stmts"]), t["n_fwd"], t["result_reg"])]
    out.append("data: " + ", ".join("%s[%d,%d]" % (d["name"], d["rows"], d["cols"]) for d in t["data"]))
    for i, s in enumerate(t["stmts"]):
        if i == 0:
            out.append("[forward]")
        if i == t["n_fwd"]:
            out.append("[gradient]")
        name = OPS.get(s["op"], "op%d" % s["op"])
        args = ", ".join(_reg(t, r) for r in s["src"])
        extra = ""
        if name in ("LOGPDF", "DLOGPDF"):
            name += DISTS[s["aux0"]] + (("/d arg %d" % s["aux1"]) if name == "DLOGPDF" else "")
        elif name == "MATMUL":
            extra = "  (%d x %d) * (%d x %d)" % (s["aux0"], s["aux1"], s["aux1"], int(s["imm"]))
        elif name in ("GATHER", "SCATTER_SET", "SCATTER_ADD"):
            extra = "  through table %d (%d indices)" % (s["aux0"], len(t["tables"][s["aux0"]]))
        elif name == "GLM":
            extra = "  family %s, aux1 = %d, parameter %r, G -> %s" % (GLM[s["aux0"]], s["aux1"], s["imm"], _reg(t, s["aux2"]))
        elif name == "FUSED":
            extra = "  program %d over %d elements" % (s["aux0"], s["aux1"])
        elif name == "AFFNORM":
            extra = "  n = %d, descriptor table %d: %s" % (s["aux1"], s["aux0"], t["tables"][s["aux0"]])
        out.append("  %-22s %s %s(%s)%s%s" % (_reg(t, s["dst"]), "+=" if s["flags"] & 1 else " =", name, args, _flags(s["flags"]), extra))
    for k, pr in enumerate(t["programs"]):
        out.append("program %d: %d elements, %d instructions (%d once per thread), reductions -> %s" % (
            k, pr["n"], len(pr["instr"]), pr["n_prologue"],
            ", ".join("%s %s" % (_reg(t, reg), "+=" if acc else "=") for reg, acc in pr["red"]) or "none"))
        for i, ins in enumerate(pr["instr"]):
            name = OPS.get(ins["op"], "op%d" % ins["op"])
            if name in ("loadv", "loads"):
                txt = "L%d = %s" % (ins["dst"], _reg(t, ins["reg"]))
            elif name == "storev":
                txt = "%s = L%d" % (_reg(t, ins["reg"]), ins["src"][0])
            elif name == "reduce":
                txt = "R%d += L%d" % (ins["aux0"], ins["src"][0])
            else:
                if name in ("LOGPDF", "DLOGPDF"):
                    name += DISTS[ins["aux0"]] + (("/d arg %d" % ins["aux1"]) if name.startswith("DLOGPDF") else "")
                txt = "L%d %s %s(%s)" % (ins["dst"], "+=" if ins["flags"] & 1 else "=", name, ", ".join("L%d" % x for x in ins["src"]))
            out.append("  %s%s" % ("| " if i < pr["n_prologue"] else "  ", txt))
    return out


if __name__ == "__main__":
    if len(sys.argv) != 2:
        sys.exit(__doc__)
    print("\n".join(lines(parse(open(sys.argv[1], "rb").read()))))
