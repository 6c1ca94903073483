"""Device checks of the stand-alone tcgen05 int8 probes under tools/ (not part of the library, so
these tests come last): the int8 UMMA instruction against a host integer product, and the sliced-integer GLM prototype
against its built-in long-double host reference at the parity bar of the suite (1e-12)."""
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


@pytest.mark.parametrize("variant", [5, 7])          # v4 and v5, 512 threads
def test_sliced_integer_glm_prototype_meets_parity(variant):
    (res,) = _run("ozaki_glm_proto", "check", 4096, 64, 3, variant)
    assert res["timeout"] == 0
    assert res["relerr_logp"] <= 1e-12 and res["relerr_grad"] <= 1e-12 and res["relerr_eta_tile0"] <= 1e-12
    assert res["meets_1e-12"] is True
