"""Statement fusion: turns runs of elementwise / reducing tape statements over one index space into a single
SMC_OP_FUSED statement carrying a small register program (include/simplemcmc_tape.h, smc_prog_instr).

The reference's generated function allocates one temporary array per statement ("every vectorised
statement allocates a fresh temporary"); with C chains in lockstep that would be [n][C] doubles per statement.  A fused
group keeps every intermediate in a per-thread register file instead: arrays are read once (LOADV), scalars once per
thread (LOADS), reductions (`sum`, `dot`, vectorised logpdf -- src/distribs.jl:97-105) accumulate per thread and are
combined in fixed order.  Array values another group needs are recomputed there (rematerialised) when their producers are
pure elementwise statements, so forward temporaries such as `resid` never touch HBM.
"""
import struct

from . import lowering as L

OP_FUSED = 49
P_LOADV, P_LOADS, P_STOREV, P_REDUCE = 100, 101, 102, 103
MAX_LOCALS = 48
MAX_REDUCE = 8
MAX_INSTR = 240

_ELEMENTWISE = {L.OP[k] for k in ("COPY", "NEG", "EXP", "LOG", "SIN", "COS", "ABS", "SIGN", "DIGAMMA", "ZERO", "SQRT", "ADD",
                                  "SUB", "MUL", "DIV", "POW", "MAX", "MIN", "GT", "LT", "GE", "LE", "EQ", "LOGPDF", "DLOGPDF")}


def index_space(st):
    n = st.dst.numel
    for s in st.srcs:
        n = max(n, s.numel)
    if st.op == L.OP["LOGPDF"]:
        n = max(st.aux1, 1)
    return n


def fusable(st):
    return st.op in _ELEMENTWISE and st.dst.kind in (L.REG_CHAIN, L.REG_GRAD)


def is_reduce(st):
    return fusable(st) and st.dst.numel == 1 and index_space(st) > 1


class Group(object):
    def __init__(self, n, fus):
        self.n, self.fusable, self.stmts = n, fus, []
        self.writes, self.reads, self.red_out = set(), set(), set()

    def add(self, st):
        self.stmts.append(st)
        self.writes.add(st.dst.vid)
        for v in st.extra_outs:
            self.writes.add(v.vid)
        for s in st.srcs:
            self.reads.add(s.vid)
        if st.flags & L.F_ACC:
            self.reads.add(st.dst.vid)
        if is_reduce(st):
            self.red_out.add(st.dst.vid)


def form_groups(stmts):
    """greedy clustering with hoisting: a statement joins the earliest group, not before any group it depends on, that
    iterates the same index space and does not hand it a reduction result"""
    groups = []
    for st in stmts:
        reads = {s.vid for s in st.srcs}
        if st.flags & L.F_ACC:
            reads.add(st.dst.vid)
        writes = {st.dst.vid} | {v.vid for v in st.extra_outs}
        gmin = 0
        for gi, g in enumerate(groups):
            if (g.writes & reads) or (g.writes & writes) or (g.reads & writes):
                gmin = gi
        fus = fusable(st)
        n = index_space(st)
        placed = False
        if fus:
            nred = 1 if is_reduce(st) else 0
            for gi in range(gmin, len(groups)):
                g = groups[gi]
                if not g.fusable or g.n != n:
                    continue
                # `acc += sum(...)` into a scalar the group already reduces into (the priors of several parameter vectors of one
                # length, src/parsing.jl:103-109 after coalesce_accumulator): both sums share one reduction slot
                same_acc = nred == 1 and (st.flags & L.F_ACC) and st.dst.vid in g.red_out
                if gi == gmin and (g.red_out & (reads - {st.dst.vid} if same_acc else reads)):   # needs a reduction finished by that very group
                    continue
                if (g.red_out & writes) and not same_acc:        # two writers of one scalar: keep them in separate groups
                    continue
                if len(g.stmts) >= 60 or sum(1 for s in g.stmts if is_reduce(s)) + nred > MAX_REDUCE:
                    continue
                # later groups must not conflict with moving st up into g
                ok = True
                for h in groups[gi + 1:]:
                    if (h.writes & reads) or (h.writes & writes) or (h.reads & writes):
                        ok = False
                        break
                if ok:
                    g.add(st)
                    placed = True
                    break
        if not placed:
            g = Group(n, fus)
            g.add(st)
            groups.append(g)
    return groups


def rematerialise(groups, producers):
    """clone cheap elementwise producers of external array inputs into the consuming group"""
    for g in groups:
        if not g.fusable or g.n <= 1:
            continue
        budget = 24
        changed = True
        while changed and budget > 0:
            changed = False
            produced = set()
            for st in g.stmts:
                ext = [s for s in st.srcs if s.numel == g.n and s.varying and s.vid not in produced]
                for v in ext:
                    p = producers.get(v.vid)
                    if p is None or len(p) != 1:
                        continue
                    p = p[0]
                    if not fusable(p) or is_reduce(p) or (p.flags & L.F_ACC) or index_space(p) != g.n or p.dst.kind != L.REG_CHAIN:
                        continue
                    if any(len(producers.get(s.vid, [])) > 1 for s in p.srcs):     # reads a multi-writer accumulator
                        continue
                    clone = L.Stmt(p.op, p.dst, p.srcs, flags=p.flags, aux0=p.aux0, aux1=p.aux1, aux2=p.aux2, imm=p.imm)
                    clone.clone = True
                    g.stmts.insert(g.stmts.index(st), clone)
                    g.writes.add(p.dst.vid)
                    for s in p.srcs:
                        g.reads.add(s.vid)
                    budget -= 1
                    changed = True
                    break
                if changed:
                    break
                produced.add(st.dst.vid)
    return groups


class Program(object):
    def __init__(self, n):
        self.n = n
        self.prologue, self.body, self.reductions = [], [], []   # instr tuples; reductions: (vid, acc)

    def touched(self):
        out = []
        for ins in self.prologue + self.body:
            if ins[6] is not None and ins[6] not in out:
                out.append(ins[6])
        for vid, _ in self.reductions:
            if vid not in out:
                out.append(vid)
        return out

    def pack(self, regmap):
        out = [struct.pack("<4I", len(self.prologue) + len(self.body), len(self.prologue), len(self.reductions), self.n)]
        for (op, dst, srcs, flags, aux0, aux1, vid) in self.prologue + self.body:
            s = list(srcs) + [0] * (3 - len(srcs))
            if op == L.OP["LOGPDF"]:
                aux1 = 0        # the statement's aux1 is its index-space length; the program header carries it as `n`
            out.append(struct.pack("<HBB3BBHHI", op, dst, len(srcs), s[0], s[1], s[2], flags, aux0, aux1,
                                   regmap[vid] if vid is not None else 0))
        for vid, acc in self.reductions:
            out.append(struct.pack("<2I", regmap[vid], acc))
        return b"".join(out)


def external_inputs(g):
    """vids a group reads before (or without) writing them itself"""
    written, ext = set(), set()
    for st in g.stmts:
        for s in st.srcs:
            if s.vid not in written:
                ext.add(s.vid)
        if (st.flags & L.F_ACC) and st.dst.vid not in written:
            ext.add(st.dst.vid)
        for v in st.extra_outs:
            written.add(v.vid)
        if not is_reduce(st):
            written.add(st.dst.vid)
    return ext


def build_program(g, outside):
    """group -> Program, or None when it does not fit the interpreter's limits.
    Locals: scalars (and everything in an n == 1 group) are loaded once per thread in the prologue and keep their slot;
    array values are loaded / computed per element in the body and their slots are recycled by liveness."""
    prog = Program(g.n)
    scalar_group = (g.n == 1)
    last_access, last_write = {}, {}
    for i, st in enumerate(g.stmts):
        for s in st.srcs:
            last_access[s.vid] = i
        if not is_reduce(st):
            last_access[st.dst.vid] = i
            last_write[st.dst.vid] = i
    local, pinned = {}, set()
    free = list(range(MAX_LOCALS - 1, -1, -1))

    def alloc(key):
        if not free:
            raise OverflowError
        local[key] = free.pop()
        return local[key]

    if not scalar_group:
        # scalar inputs: loaded once per thread (prologue) into slots no per-element value may ever reuse
        for st in g.stmts:
            for v in st.srcs:
                if v.numel == 1 and v.vid not in local:
                    if len(local) >= MAX_LOCALS - 8:
                        return None
                    prog.prologue.append((P_LOADS, alloc(v.vid), [], 0, 0, 0, v.vid))
                    pinned.add(v.vid)

    def load(v):
        if v.vid in local:
            return local[v.vid]
        slot = alloc(v.vid)
        prog.body.append((P_LOADS if scalar_group else P_LOADV, slot, [], 0, 0, 0, v.vid))
        return slot

    try:
        for i, st in enumerate(g.stmts):
            srcs = [load(s) for s in st.srcs]
            acc = 1 if (st.flags & L.F_ACC) else 0
            if is_reduce(st):
                slot = [k for k, (vid, _) in enumerate(prog.reductions) if vid == st.dst.vid]
                if slot and not acc:
                    return None                      # (a second `=` into one scalar: form_groups never builds this)
                if not slot:
                    if len(prog.reductions) >= MAX_REDUCE:
                        return None
                    slot = [len(prog.reductions)]
                    prog.reductions.append((st.dst.vid, acc))
                tmp = alloc(("tmp", i))
                prog.body.append((st.op, tmp, srcs, 0, st.aux0, st.aux1, None))
                prog.body.append((P_REDUCE, 0, [tmp], 0, slot[0], 0, None))
                free.append(local.pop(("tmp", i)))
            else:
                if acc and st.dst.vid not in local:
                    load(st.dst)
                dst = local[st.dst.vid] if st.dst.vid in local else alloc(st.dst.vid)
                if st.dst.vid in pinned:
                    return None                      # would overwrite a prologue scalar inside the element loop
                prog.body.append((st.op, dst, srcs, acc, st.aux0, st.aux1, None))
                if last_write[st.dst.vid] == i and st.dst.vid in outside and not getattr(st, "clone", False):
                    prog.body.append((P_STOREV, 0, [dst], 0, 0, 0, st.dst.vid))
            for v in list(st.srcs) + [st.dst]:
                if v.vid in local and v.vid not in pinned and last_access.get(v.vid, -1) <= i:
                    free.append(local.pop(v.vid))
    except OverflowError:
        return None
    if len(prog.prologue) + len(prog.body) > MAX_INSTR:
        return None
    return prog


def hand_over_reductions(gf, gb, producers, body_f, result_vid, always):
    """An evaluation WITH gradient runs the forward groups and then the backward groups; a backward group usually recomputes
    (rematerialise) the forward values of its index space, so the data of that index space is walked twice -- once for
    `acc += sum(logpdf...)`, once for the adjoints.  Where a backward group B already holds clones of every elementwise
    statement of a forward group F, F's reductions into the result accumulator are handed to B and F is marked
    SMC_F_FWD_ONLY: the library runs it only when no gradient is asked for (include/simplemcmc_tape.h).  Only `+=`
    reductions move (the accumulator's first, plain `=` writer has to stay first), and only when F stores nothing another
    statement reads.  -> set of id(F) of the groups that became forward-only"""
    fwd_only = set()
    fwd_ids = {id(s) for s in body_f}
    for F in gf:
        if not F.fusable or F.n <= 1 or not gb:
            continue
        reds = [s for s in F.stmts if is_reduce(s)]
        plain = [s for s in F.stmts if not is_reduce(s)]
        if not reds or any(not (r.flags & L.F_ACC) or r.dst.vid != result_vid for r in reds):
            continue
        others = set(always)
        for g in gf + gb:
            if g is not F:
                others |= external_inputs(g)
        if any(p.dst.vid in others for p in plain):
            continue                                  # F materialises a value somebody else reads: it has to run
        own = {p.dst.vid for p in plain}
        for B in gb:
            if not B.fusable or B.n != F.n:
                continue
            have = {}
            for i, st in enumerate(B.stmts):
                if not is_reduce(st):
                    have[st.dst.vid] = i
            if any(p.dst.vid not in have or not getattr(B.stmts[have[p.dst.vid]], "clone", False) for p in plain):
                continue

            def available(v):
                if v.vid in own:
                    return True
                if v.kind in (L.REG_IMM, L.REG_DATA, L.REG_HDR, L.REG_PARAM):
                    return True
                pr = producers.get(v.vid, [])                 # a forward value computed outside F, once
                return v.kind == L.REG_CHAIN and len(pr) == 1 and id(pr[0]) in fwd_ids and pr[0] not in F.stmts
            if not all(available(v) for r in reds for v in r.srcs):
                continue
            trial = Group(B.n, True)
            stmts = list(B.stmts)
            for r in reds:
                at = max([have_at for have_at in (_index_of_writer(stmts, v.vid) for v in r.srcs) if have_at is not None] + [-1]) + 1
                moved = L.Stmt(r.op, r.dst, r.srcs, flags=r.flags, aux0=r.aux0, aux1=r.aux1, aux2=r.aux2, imm=r.imm)
                stmts.insert(at, moved)
            for st in stmts:
                trial.add(st)
            outside = set(always)
            for g in gf + gb:
                if g is not B:
                    outside |= external_inputs(g)
            if build_program(trial, outside) is None:
                continue
            if build_program(F, set(always)) is None:         # (F has to stay one statement for the flag to cover all of it)
                continue
            B.stmts, B.writes, B.reads, B.red_out = trial.stmts, trial.writes, trial.reads, trial.red_out
            fwd_only.add(id(F))
            break
    return fwd_only


def _index_of_writer(stmts, vid):
    at = None
    for i, st in enumerate(stmts):
        if not is_reduce(st) and st.dst.vid == vid:
            at = i
    return at


def coalesce_accumulator(fwd, bwd, result):
    """`__acc = __acc + logpdf...` (src/parsing.jl:103-109) leaves a chain of scalar additions a1 = t1 + t2, a2 = a1 + t3, ...
    whose operands are each produced once and consumed once.  Instead of one statement (one launch) per addition the
    later producer accumulates straight into the earlier operand's register (`x += value`) and the sum becomes another
    name for x.  Same additions in the same order, so the value is bit-identical.  -> (fwd, bwd, result)"""
    fwd, bwd = list(fwd), list(bwd)
    changed = True
    while changed:
        changed = False
        uses = {}
        for st in fwd + bwd:
            for v in st.srcs:
                uses[v.vid] = uses.get(v.vid, 0) + 1
        prod = {}
        for i, st in enumerate(fwd):
            prod.setdefault(st.dst.vid, []).append(i)
            for v in st.extra_outs:
                prod.setdefault(v.vid, []).append(i)
        for i, st in enumerate(fwd):
            if st.op != L.OP["ADD"] or st.flags or st.dst.numel != 1 or st.dst.kind != L.REG_CHAIN:
                continue
            x, y = st.srcs
            if x is y or any(v.numel != 1 or v.kind != L.REG_CHAIN or uses.get(v.vid, 0) != 1 for v in (x, y)):
                continue
            px, py = prod.get(x.vid, []), prod.get(y.vid, [])
            if len(py) != 1 or not px or max(px) >= py[0] or py[0] >= i:
                continue                      # y must have one producer, after every writer of x and before the sum
            p = fwd[py[0]]
            if (p.flags & L.F_ACC) or p.dst is not y or not (fusable(p) or p.op in (L.OP["GLM"], L.OP["AFFNORM"])):
                continue
            q = L.Stmt(p.op, x, p.srcs, flags=p.flags | L.F_ACC, aux0=p.aux0, aux1=p.aux1, aux2=p.aux2, imm=p.imm, desc=p.desc)
            fwd[py[0]] = q
            t = st.dst
            del fwd[i]
            for other in fwd + bwd:           # the sum is now another name for x
                if any(v is t for v in other.srcs):
                    other.srcs = [x if v is t else v for v in other.srcs]
            if result is t:
                result = x
            changed = True
            break
    return fwd, bwd, result


def fuse(lowered):
    """-> (fwd, bwd, programs): statement lists in which runs of elementwise statements are FUSED statements"""
    fwd0 = [s.copy() for s in lowered.fwd]
    bwd0 = [s.copy() for s in lowered.bwd]
    fwd0, bwd0, lowered.tape_result = coalesce_accumulator(fwd0, bwd0, lowered.result)
    # The library keeps the result register zero between evaluations (tape version 4, csrc/smc_core.cu: smc_clear_result), so
    # its first writer adds to it like every later one (0 + x is exact): no contribution to the accumulator is tied to a place
    # in the statement order any more, which is what lets hand_over_reductions move them
    for st in fwd0:
        if st.dst is lowered.tape_result:
            if not (st.flags & L.F_ACC) and st.dst.numel == 1 and (fusable(st) or st.op in (L.OP["GLM"], L.OP["AFFNORM"])):
                st.flags |= L.F_ACC
            break
    body_f = [s for s in fwd0 if s.dst.kind != L.REG_HDR]
    body_b = [s for s in bwd0 if s.dst.kind != L.REG_HDR]
    out_f = [s for s in fwd0 if s.dst.kind == L.REG_HDR]
    out_b = [s for s in bwd0 if s.dst.kind == L.REG_HDR]
    producers = {}
    for s in body_f + body_b:
        producers.setdefault(s.dst.vid, []).append(s)
    gf = rematerialise(form_groups(body_f), producers)
    gb = rematerialise(form_groups(body_b), producers)
    always = {lowered.tape_result.vid} | {v.vid for v in lowered.vals if v.kind == L.REG_GRAD}
    # a statement whose value every consumer now recomputes for itself (rematerialisation) is dead: drop it
    changed = True
    while changed:
        changed = False
        groups_now = gf + gb
        ext = [external_inputs(g) for g in groups_now]
        for gi, g in enumerate(groups_now):
            if any(st.op in (L.OP["GLM"], L.OP["AFFNORM"]) for st in g.stmts):
                continue
            others = set().union(*[e for hj, e in enumerate(ext) if hj != gi]) if len(ext) > 1 else set()
            if any(st.dst.vid in always or st.dst.vid in others for st in g.stmts):
                continue
            (gf if g in gf else gb).remove(g)
            changed = True
            break
    fwd_only = hand_over_reductions(gf, gb, producers, body_f, lowered.tape_result.vid, always)
    all_groups = gf + gb
    ext = [external_inputs(g) for g in all_groups]
    programs = []
    for gi, g in enumerate(all_groups):
        target = out_f if gi < len(gf) else out_b
        if not g.fusable or (len(g.stmts) < 2 and not is_reduce(g.stmts[0])):
            target.extend(g.stmts)           # (a lone reduction still becomes a program: one launch instead of two)
            continue
        outside = set(always)
        for hj, e in enumerate(ext):
            if hj != gi:
                outside |= e
        prog = build_program(g, outside)
        dst = None
        if prog is not None:
            written = {st.dst.vid for st in g.stmts}
            for vid in prog.touched():
                v = lowered.vals[vid]
                if v.kind in (L.REG_CHAIN, L.REG_GRAD) and vid in written and vid in outside:
                    dst = v
                    break
        if prog is None or dst is None:
            if id(g) in fwd_only:                   # (hand_over_reductions built this very program: cannot happen)
                raise AssertionError("a forward-only group did not become a program")
            target.extend(s for s in g.stmts if not getattr(s, "clone", False))
            continue
        st = L.Stmt(OP_FUSED, dst, [], flags=L.F_FWD_ONLY if id(g) in fwd_only else 0, aux0=len(programs), aux1=g.n)
        st.touched = [lowered.vals[v] for v in prog.touched()]
        programs.append(prog)
        target.append(st)
    return out_f, out_b, programs
