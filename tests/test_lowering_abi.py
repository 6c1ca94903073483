"""Host logic of the product (no GPU needed): model -> tape lowering, tape serialisation against the C header,
GLM fusion, and that the C-ABI library loads and exports every symbol include/simplemcmc.h declares."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest

from smc_pkg import load_package
from util import HIERARCHICAL, LINEAR, LOGISTIC, ORNSTEIN, SLEEPSTUDY2, hierarchical_data, linear_data, logistic_data, ou_data, sleepstudy_data

load_package()
from simplemcmc_jl_b200 import capi, lowering  # noqa: E402
from simplemcmc_jl_b200.lowering import Lowering, ModelError  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_enums(path):
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return {m.group(1): int(m.group(2)) for m in re.finditer(r"\b(SMC_[A-Z0-9_]+)\s*=\s*(\d+)", txt)}


def test_python_constants_match_the_c_header():
    e = _header_enums(os.path.join(ROOT, "include", "simplemcmc_tape.h"))
    for name, val in lowering.OP.items():
        assert e["SMC_OP_" + name] == val, name
    for name, val in lowering.DIST.items():
        assert e["SMC_DIST_" + name.upper()] == val, name
    assert (e["SMC_REG_IMM"], e["SMC_REG_DATA"], e["SMC_REG_HDR"], e["SMC_REG_CHAIN"], e["SMC_REG_PARAM"], e["SMC_REG_GRAD"]) == (0, 1, 2, 3, 4, 5)
    assert e["SMC_GLM_NORMAL"] == lowering.GLM_NORMAL and e["SMC_GLM_BERNOULLI"] == lowering.GLM_BERNOULLI
    txt = open(os.path.join(ROOT, "include", "simplemcmc_tape.h")).read()
    assert "#define SMC_TAPE_VERSION %du" % lowering.TAPE_VERSION in txt


def test_library_loads_and_exports_the_whole_abi():
    hdr = open(os.path.join(ROOT, "include", "simplemcmc.h")).read()
    declared = set(re.findall(r"\b(smc_[a-z_0-9]+)\s*\(", hdr))
    lib = capi.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(capi.EXPORTS)
    assert lib.smc_abi_version() == 2
    assert ctypes.sizeof(capi.SamplerConfig) == 80 and ctypes.sizeof(capi.RunOutput) == 120


def test_ctx_create_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.SMCError) as e:
        capi.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_linear_lowering_fuses_the_glm():
    X, Y, _ = linear_data()
    low = Lowering(LINEAR, [("beta", np.zeros(10))], dict(X=X, Y=Y))
    fw, bw = low.statements()
    assert low.n_fused == 1 and [s.op for s in fw].count(lowering.OP["GLM"]) == 1
    assert not any(s.dst.numel == 1000 for s in fw + bw)          # no length-N per-chain register survives
    low.fuse_statements = False
    tape, bound = low.serialise()
    magic, ver, nreg, nst, npar, ndata, nidx, res, flags, nfwd, nprog, _ = struct.unpack_from("<12I", tape)
    assert (magic, ver, npar, ndata, nfwd, nprog) == (0x54434D53, lowering.TAPE_VERSION, 10, 2, len(fw), 0)
    assert len(tape) == 48 + 32 * nreg + 56 * nst + 64 * ndata
    low.fuse_statements = True
    tape2, _ = low.serialise()
    nst2, nprog2 = struct.unpack_from("<12I", tape2)[3], struct.unpack_from("<12I", tape2)[10]
    assert nprog2 >= 1 and nst2 < nst                          # the prior / adjoint statements collapse into programs
    assert sorted(n for n, _ in bound) == ["X", "Y"]
    unf = Lowering(LINEAR, [("beta", np.zeros(10))], dict(X=X, Y=Y), fuse=False)
    assert unf.n_fused == 0 and any(s.op == lowering.OP["MATMUL"] and s.flags & lowering.F_TRANS_A for s in unf.bwd)


def test_logistic_lowering_fuses_with_the_reference_sign():
    X, Y, _ = logistic_data()
    low = Lowering(LOGISTIC, [("vars", np.zeros(10))], dict(X=X, Y=Y))
    glm = [s for s in low.fwd if s.op == lowering.OP["GLM"]]
    assert len(glm) == 1 and glm[0].aux0 == lowering.GLM_BERNOULLI and glm[0].aux1 == 0    # prob = 1/(1+exp(+eta))
    low2 = Lowering("v ~ Normal(0,1)\nprob = 1. / (1. + exp(-(X * v)))\nY ~ Bernoulli(prob)", [("v", np.zeros(10))], dict(X=X, Y=Y))
    glm2 = [s for s in low2.fwd if s.op == lowering.OP["GLM"]]
    assert len(glm2) == 1 and glm2[0].aux1 == 1


def test_example_models_lower():
    x = ou_data(200)
    low = Lowering(ORNSTEIN, [("tau", 20.0), ("mu", 10.0), ("sigma", 0.1)], dict(x=x))
    assert low.bsize == 3 and len(low.idx_tables) == 2
    hdr = [s for s in low.fwd if s.dst.kind == lowering.REG_HDR]
    assert len(hdr) == 2                                           # x[2:end], x[1:end-1] are hoisted (parsing.jl:432-442)
    data, (N, D, L) = hierarchical_data()
    low = Lowering(HIERARCHICAL, [("mu", np.zeros((1, L))), ("sigma", np.ones((1, L))), ("beta", np.zeros((D, L)))], data)
    assert low.bsize == 30 and [p.start for p in low.pars] == [0, 5, 10]
    assert any(s.op == lowering.OP["SCATTER_SET"] for s in low.bwd)
    low = Lowering(HIERARCHICAL, [("mu", np.zeros((1, L))), ("sigma", np.ones((1, L))), ("beta", np.zeros((D, L)))], data, ref_adjoint="add")
    assert any(s.op == lowering.OP["SCATTER_ADD"] for s in low.bwd)
    data, nsub = sleepstudy_data()
    low = Lowering(SLEEPSTUDY2, [("intercept", 250.0), ("slope", 10.0), ("rand_int", np.zeros(nsub)), ("rand_slope", np.zeros(nsub))], data)
    assert low.bsize == 38


def test_statement_fusion_of_the_ornstein_tape():
    """cfg5: the 46 statements become a handful of register programs; no length-T per-chain register is stored"""
    from simplemcmc_jl_b200 import fusion
    x = ou_data(500)
    low = Lowering(ORNSTEIN, [("tau", 20.0), ("mu", 10.0), ("sigma", 0.1)], dict(x=x))
    low.serialise()
    stmts = low.tape_fwd + low.tape_bwd
    assert len(stmts) <= 12 and len(low.programs) >= 4
    stored = {ins[6] for pr in low.programs for ins in pr.body if ins[0] == fusion.P_STOREV}
    assert all(low.vals[v].numel == 1 for v in stored)          # resid & co. are rematerialised, never written to HBM
    for pr in low.programs:                                     # prologue scalars never share a slot with body values
        pinned = {ins[1] for ins in pr.prologue}
        written = {ins[1] for ins in pr.body if ins[0] not in (fusion.P_STOREV, fusion.P_REDUCE)}
        assert not (pinned & written)


def test_tape_identities_and_dead_statement_elimination():
    """the reference's TODO.md:4-9 ("+0, *1", constant folding) done at lowering time; values are unchanged (exact identities)"""
    low = Lowering("y = (x + 0) * 1\nz = y / 1 - 0\nz ~ Normal(0 + 0, 2 * 0.5)", [("x", np.zeros(3))], {})
    assert [s.op for s in low.fwd] == [lowering.OP["LOGPDF"]]            # only the density is left
    assert low.fwd[0].srcs[0].imm == 0.0 and low.fwd[0].srcs[1].imm == 1.0 and low.fwd[0].srcs[2].kind == lowering.REG_PARAM
    # sleepstudy: `intercept + submatrix*rand_int` is recomputed inside the programs that use it, so the stand-alone
    # statement that would store it as a [180][C] array is dropped
    data, nsub = sleepstudy_data()
    low = Lowering(SLEEPSTUDY2, [("intercept", 250.0), ("slope", 10.0), ("rand_int", np.zeros(nsub)), ("rand_slope", np.zeros(nsub))], data)
    low.serialise()
    plain = [s for s in low.tape_fwd if s.op not in (lowering.OP["MATMUL"], 49)]
    assert plain == [], plain
    assert sum(1 for s in low.tape_fwd + low.tape_bwd if s.op == lowering.OP["MATMUL"]) == 4


def test_linear_tape_is_three_statements_after_the_host_rewrites():
    """accumulator coalescing + literal seed + adjoint names (fusion.py, lowering.py): prior program, GLM (accumulating),
    gradient program -- and the result register is the one both forward statements write"""
    X, Y, _ = linear_data()
    low = Lowering(LINEAR, [("beta", np.zeros(10))], dict(X=X, Y=Y))
    low.serialise()
    assert [s.op for s in low.tape_fwd] == [49, lowering.OP["GLM"]] and [s.op for s in low.tape_bwd] == [49]
    assert low.tape_fwd[0].dst is low.tape_result and low.tape_fwd[1].dst is low.tape_result
    assert low.tape_fwd[1].flags & lowering.F_ACC and not (low.tape_fwd[0].flags & lowering.F_ACC)
    assert low.tape_bwd[0].dst.kind == lowering.REG_GRAD


def test_sufficient_statistics_rewrite_of_the_ornstein_likelihood():
    """opt-in: the length-T Normal density becomes scalar arithmetic on five header dot products (x.x, x.x', sum x, ...)"""
    x = ou_data(500)
    low = Lowering(ORNSTEIN, [("tau", 20.0), ("mu", 10.0), ("sigma", 0.1)], dict(x=x), sufficient_stats=True)
    low.serialise()
    hdr = [s for s in low.tape_fwd if s.dst.kind == lowering.REG_HDR]
    assert sum(1 for s in hdr if s.dst.numel == 1) == 5                       # G = [[x1.x1, x1.x0, sum x1], [., x0.x0, sum x0], [., ., n]]
    body = [s for s in low.tape_fwd + low.tape_bwd if s.dst.kind != lowering.REG_HDR]
    assert all(s.op == 49 for s in body) and all(pr.n == 1 for pr in low.programs)   # only scalar programs are left
    plain = Lowering(ORNSTEIN, [("tau", 20.0), ("mu", 10.0), ("sigma", 0.1)], dict(x=x))
    plain.serialise()
    assert any(pr.n == 499 for pr in plain.programs)                            # the default keeps the element-wise tape


def test_programs_pack_for_index_spaces_beyond_16_bits():
    """cfg5 at T = 1e6: a program's index-space length lives in its header, not in a 16-bit instruction field"""
    import struct
    x = ou_data(100_000)
    low = Lowering(ORNSTEIN, [("tau", 20.0), ("mu", 10.0), ("sigma", 0.1)], dict(x=x), affnorm=False)
    tape, bound = low.serialise()
    assert any(pr.n == 99_999 for pr in low.programs)
    blob = low.programs[[pr.n for pr in low.programs].index(99_999)].pack({v.vid: v.reg for v in low.vals if v.reg is not None})
    assert struct.unpack("<4I", blob[:16])[3] == 99_999


def test_ornstein_residual_becomes_one_affnorm_statement():
    """examples/ornstein.jl:24-26: `resid = x[2:end] - x[1:end-1] * fac - mu * (1. - fac); resid ~ Normal(0, sigma)` is one
    AFFNORM statement whose two array terms are views (offsets 1 and 0) of the ONE bound vector x -- no header gathers, no
    element-wise temporaries; what is left around it are two scalar programs"""
    import struct
    x = ou_data(5000)
    low = Lowering(ORNSTEIN, [("tau", 20.0), ("mu", 10.0), ("sigma", 0.1)], dict(x=x))
    assert low.n_affnorm == 1
    tape, bound = low.serialise()
    assert [b[0] for b in bound] == ["x"]
    aff = [s for s in low.tape_fwd if s.op == lowering.OP["AFFNORM"]]
    assert len(aff) == 1 and aff[0].flags & lowering.F_ACC          # accumulates into __acc after the three priors
    d = aff[0].desc
    assert d.n == 4999 and [(k, sg, off) for (k, sg, b, off) in d.terms] == [(0, 1.0, 1), (0, -1.0, 0), (1, -1.0, 0)]
    assert d.terms[0][2] is d.terms[1][2] and d.terms[0][2].kind == lowering.REG_DATA
    assert d.outs[0] is None and d.outs[1] is not None                # mu = 0 literal; sigma is a parameter
    assert d.outs[2] is None and d.outs[3] is not None and d.outs[4] is not None   # coefficient 1, -fac, -mu (1 - fac)
    ops = [s.op for s in low.tape_fwd + low.tape_bwd]
    assert ops == [49, lowering.OP["AFFNORM"], 49] and all(pr.n == 1 for pr in low.programs)
    # the descriptor travels as the last index table of the blob
    hdr = struct.unpack("<12I", tape[:48])
    assert hdr[6] == 1 and len(tape) < 2000                        # the two range tables of the gathers are not shipped
    # below the size threshold, or when a temporary is used twice, the element-wise tape stays
    assert Lowering(ORNSTEIN, [("tau", 20.0), ("mu", 10.0), ("sigma", 0.1)], dict(x=ou_data(500))).n_affnorm == 0
    twice = ORNSTEIN + "\nsum(resid) ~ Normal(0, 100)"
    assert Lowering(twice, [("tau", 20.0), ("mu", 10.0), ("sigma", 0.1)], dict(x=x)).n_affnorm == 0


def test_affnorm_linearisation_keeps_the_reference_roundings():
    """caterpillar trees of + / - / unary minus are chains (IEEE addition commutes, negation is exact); a node with two
    non-leaf children is not"""
    d = dict(a=np.ones(4096), b=np.ones(4096), c=np.ones(4096))
    def terms(expr):
        low = Lowering("r = %s\nr ~ Normal(m, 2.0)" % expr, [("s", 1.5), ("m", 0.5)], d)
        return None if low.n_affnorm == 0 else [(k, sg) for (k, sg, _, _) in [s for s in low.fwd if s.op == lowering.OP["AFFNORM"]][0].desc.terms]
    assert terms("a * s - b") == [(0, 1.0), (0, -1.0)]
    assert terms("s - (a - b * s)") == [(0, -1.0), (0, 1.0), (1, 1.0)]            # -(a - b s) + s
    assert terms("-(a + s) + c * s") == [(0, -1.0), (1, -1.0), (0, 1.0)]
    assert terms("(a - s) + (b - c)") == [(0, 1.0), (1, -1.0), (0, 1.0)]          # b - c is parameter independent: a header array
    assert terms("(a - s) + (b * s - c)") is None                                # two chains meet: the rounding differs
    assert terms("a / s - b") is None                                            # division is not a term
    assert terms("(a * s) * s - b") is None                                      # (a s) s: two roundings in one term


def test_reference_error_cases():
    with pytest.raises(ModelError):
        Lowering("sum(logpdfBernoulli(1, x))", [("x", np.array([0.0]))], {})
    with pytest.raises(ModelError):
        Lowering("y = 2 * x\nz ~ Normal(0, 1)", [("x", 1.0)], dict(z=np.ones(3)))
    with pytest.raises(ModelError):
        Lowering("x ~ Normal(0, 1)", [], {})
    with pytest.raises(ModelError):
        Lowering("x ~ Normal(w, 1)", [("x", 1.0)], {})             # undefined external variable
    with pytest.raises(ModelError):
        Lowering("y = x .* z\ny ~ Normal(0, 1)", [("x", np.zeros(3))], dict(z=np.ones(4)))   # dimension mismatch


def test_parameter_packing_is_keyword_order_column_major():
    """setInit!, src/parsing.jl:303-334"""
    m = np.arange(6.0).reshape(2, 3)
    low = Lowering("a ~ Normal(0,1)\nb ~ Normal(0,1)\nc ~ Normal(0,1)", [("b", m), ("a", 1.5), ("c", np.array([7.0, 8.0]))], {})
    assert low.init == [0.0, 3.0, 1.0, 4.0, 2.0, 5.0, 1.5, 7.0, 8.0]
    assert [(p.sym, p.start, p.count) for p in low.pars] == [("b", 0, 6), ("a", 6, 1), ("c", 7, 2)]


def _programs(model, init, data, **kw):
    from simplemcmc_jl_b200 import fusion
    from simplemcmc_jl_b200.lowering import Lowering
    low = Lowering(model, init, data, **kw)
    _, _, progs = fusion.fuse(low)
    regmap = {v.vid: i for i, v in enumerate(low.vals)}
    return [p.pack(regmap) for p in progs]


def test_register_program_planner_forms_chains():
    """the library's planner (host code, reached through the smc_vm_plan_describe test hook): hand-over values stay in
    registers.  examples/ornstein.jl's residual `x[2:end] - x[1:end-1]*fac - mu*(1-fac)` is ONE chain micro-op
    (multiply, reversed subtract, subtract) feeding the Normal density; sleepstudy's forward program is three micro-ops,
    the last one a chain that ends in the Normal density stage."""
    x = ou_data(1000, 5)
    plans = [capi.plan_describe(p) for p in _programs(ORNSTEIN, [("tau", 20.0), ("mu", 10.0), ("sigma", 0.1)], dict(x=x))]
    fwd = [pl for pl in plans if any(l.startswith("op32(dist0") and "->R0" in l for l in pl)]   # the likelihood's program
    # (two programs hold the likelihood: the forward-only one, and the gradient program that takes its reduction over)
    assert len(fwd) == 2 and len(fwd[0]) == 3, fwd                      # chain, density + reduction, summary line
    assert fwd[0][0].startswith("chain(4,3,2) A0 S0 A1 S1 ->L0"), fwd[0]
    assert fwd[0][1].startswith("op32(dist0,arg0) S2 S3 L0 ->R0"), fwd[0]
    data, nsub = sleepstudy_data()
    init = [("intercept", 250.0), ("slope", 10.0), ("rand_int", np.zeros(nsub)), ("rand_slope", np.zeros(nsub))]
    plans = [capi.plan_describe(p) for p in _programs(SLEEPSTUDY2, init, data)]
    big = [pl for pl in plans if any(l.startswith("chain(2,8,0)") for l in pl)]
    assert len(big) == 1 and len(big[0]) == 4, plans
    assert big[0][2].startswith("chain(2,8,0) L1 A3 S2 S3 ->R0"), big[0]   # (r - reaction) then Normal log-density, reduced
    # without chains the same program needs six micro-ops
    os.environ["SMC_VM_NO_CHAINS"] = "1"
    try:
        plain = [capi.plan_describe(p) for p in _programs(SLEEPSTUDY2, init, data)]
    finally:
        del os.environ["SMC_VM_NO_CHAINS"]
    assert max(len(pl) for pl in plain) - 1 > max(len(pl) for pl in plans) - 1
    assert sum(len(pl) for pl in plain) > sum(len(pl) for pl in plans)
