"""tests/golden/*.tape: committed device op tapes of the reference's five example models (tools/make_golden_tapes.py).
The tape format (include/simplemcmc_tape.h) is the language-neutral half of the drop-in boundary: a Julia emitter that
replaces the `eval` of src/parsing.jl:488-492 can be diffed against these bytes.  Here: the Python front end reproduces
them byte for byte, every blob parses by the header's own layout, and (when the library is built) compiles."""
import os
import struct
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from smc_pkg import load_package  # noqa: E402

load_package()
import make_golden_tapes as G  # noqa: E402

CASES = list(G.cases())


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_front_end_reproduces_the_golden_tape(case):
    name, model, init, data, kw = case
    tape, bound = G.build(model, init, data, kw)
    want = open(os.path.join(ROOT, "tests", "golden", name + ".tape"), "rb").read()
    assert tape == want, "%s: the emitted tape differs from tests/golden/%s.tape (first difference at byte %d)" % (
        name, name, next((i for i, (a, b) in enumerate(zip(tape, want)) if a != b), min(len(tape), len(want))))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_golden_tape_parses_by_the_header_layout(case):
    """walk the blob exactly as include/simplemcmc_tape.h lays it out; nothing may be left over"""
    name = case[0]
    blob = open(os.path.join(ROOT, "tests", "golden", name + ".tape"), "rb").read()
    magic, version, n_regs, n_stmts, n_params, n_data, n_idx, result_reg, flags, n_fwd, n_prog, _ = struct.unpack("<12I", blob[:48])
    assert magic == 0x54434D53 and version == 4 and result_reg < n_regs and n_fwd <= n_stmts
    p = 48
    kinds = []
    for _ in range(n_regs):
        kind, rows, cols, ndim, aux, imm = struct.unpack("<4Iqd", blob[p:p + 32])
        assert kind <= 5 and ndim <= 2
        kinds.append(kind)
        p += 32
    ops = []
    for _ in range(n_stmts):
        op, fl, dst, nsrc, s0, s1, s2, s3, a0, a1, a2, pad, imm = struct.unpack("<4I4I4Id", blob[p:p + 56])
        assert dst < n_regs and nsrc <= 4 and all(s < n_regs for s in (s0, s1, s2, s3)[:nsrc]) and kinds[dst] in (2, 3, 5)
        if op == 49:
            assert a0 < n_prog
        if op == 50:
            assert a0 < n_idx and nsrc == 2
        ops.append(op)
        p += 56
    for _ in range(n_data):
        nm, rows, cols = struct.unpack("<56s2I", blob[p:p + 64])
        assert nm.rstrip(b"\0").decode().isidentifier()
        p += 64
    for _ in range(n_idx):
        n, _pad = struct.unpack("<2I", blob[p:p + 8])
        p += 8 + 4 * ((n + 1) & ~1)
    for _ in range(n_prog):
        n_instr, n_pro, n_red, n = struct.unpack("<4I", blob[p:p + 16])
        assert n_pro <= n_instr and n_red <= 8
        p += 16 + 16 * n_instr + 8 * n_red
    assert p == len(blob)
    if name == "ornstein_T5000_affnorm":
        assert ops.count(50) == 1
    if name in ("linear", "binomial"):
        assert ops.count(48) == 1


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tape_dump_reads_every_golden_tape(case):
    """tools/tape_dump.py (the readable form an emitter author diffs): parses to the last byte and names what the tape holds"""
    import tape_dump
    name = case[0]
    t = tape_dump.parse(open(os.path.join(ROOT, "tests", "golden", name + ".tape"), "rb").read())
    text = "\n".join(tape_dump.lines(t))
    assert t["version"] == 4 and "[forward]" in text
    if name == "sleepstudy":
        assert text.count("fwd-only") == 2 and "LOGPDFNormal" in text and "(18 x 180) * (180 x 1)" in text
    if name in ("linear", "binomial"):
        assert "GLM(" in text and ("family NORMAL" in text or "family BERNOULLI" in text)
    if name == "ornstein_T5000_affnorm":
        assert "AFFNORM(" in text
    if name == "linear_no_gradient":
        assert "[gradient]" not in text
    with pytest.raises(ValueError):
        tape_dump.parse(open(os.path.join(ROOT, "tests", "golden", name + ".tape"), "rb").read()[:-4])


@pytest.mark.gpu
def test_golden_tapes_compile_on_the_device():
    """smc_model_compile / bind / finalize accept the golden blobs (the library has no CPU fallback: a gpu test)"""
    from smc_pkg import load_package
    load_package()
    from simplemcmc_jl_b200 import capi
    ctx = capi.Context.default(0)
    for name, model, init, data, kw in CASES:
        tape, bound = G.build(model, init, data, kw)
        dev = capi.DeviceModel(ctx, tape, bound, max_chains=3)
        dev.close()
