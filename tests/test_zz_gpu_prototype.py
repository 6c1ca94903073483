"""Device checks of the stand-alone hardware probes under tools/ (not part of the library, so these
tests come last): the int8 UMMA instruction against a host integer product (the three operand arrangements csrc/glm_i8.cu
uses), and the probes whose findings shaped that kernel -- their qualitative result is asserted, so that a driver or a part
that behaves differently is noticed.  (The stand-alone sliced-integer GLM prototype of round 1 is gone: the library kernel
superseded it; its source is in the history at a699761, its measurements in profiles/ozaki_glm_proto_v*_r01.jsonl.)"""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _run(exe, *args):
    path = os.path.join(ROOT, "tools", exe)
    assert os.path.exists(path), "%s is not built: run python __graft_entry__.py" % path
    out = subprocess.run([path] + [str(a) for a in args], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    return [json.loads(line) for line in out.stdout.splitlines() if line.startswith("{")]


@pytest.mark.parametrize("mode,N", [("kk", 64), ("mk", 64), ("km", 64), ("kk", 256), ("mk", 256), ("km", 128)])
def test_int8_umma_is_bit_exact(mode, N):
    (res,) = _run("i8_umma_probe", "check", mode, N)
    assert res["timeout"] == 0 and res["mismatches"] == 0 and res["bit_exact"] is True


def test_fp64_is_starved_while_the_tensor_core_runs_int8_mmas():
    """the reason csrc/glm_i8.cu starts its tensor bursts where the epilogues have integer work only"""
    res = {(r["instruction"], r["tensor_core"].split()[0]): r for r in _run("fp64_vs_imma_probe")}
    assert all(r["status"] == 0 for r in res.values())
    assert res[("DFMA", "busy")]["cycles_per_instruction_per_scheduler"] > 3.0 * res[("DFMA", "idle")]["cycles_per_instruction_per_scheduler"]
    assert res[("IMAD", "busy")]["cycles_per_instruction_per_scheduler"] < 1.3 * res[("IMAD", "idle")]["cycles_per_instruction_per_scheduler"]


def test_graph_while_node_is_driven_by_a_kernel():
    """the device-side round loop of simpleNUTS (SMC_NUTS_DEVICE_LOOP=1) relies on it"""
    path = os.path.join(ROOT, "tools", "cond_graph_probe")
    out = subprocess.run([path], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.count("n = 11") == 3, out.stdout + out.stderr
