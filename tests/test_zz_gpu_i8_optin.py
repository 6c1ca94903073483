"""Opt-in parity test of the int8 tensor-core GLM path of the library (csrc/glm_i8.cu).  That path has had a
single device run so far (tools/i8_quick.py, profiles/glm_i8_library_quick_r01.json), so it is excluded from the default GPU suite:
    SMC_TEST_I8=1 python -m pytest tests/test_zz_gpu_i8_optin.py -m gpu
The prototype kernel it derives from is covered by test_zz_gpu_prototype.py."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("SMC_TEST_I8") != "1", reason="opt-in: SMC_TEST_I8=1")]


@pytest.mark.parametrize("rows,chains", [(5000, 70), (16384, 256), (1000, 32)])
def test_int8_glm_path_matches_the_oracle(rows, chains):
    env = dict(os.environ, SMC_GLM_I8="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "i8_parity_check.py"), str(rows), str(chains)],
                         capture_output=True, text=True, timeout=900, env=env)
    lines = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    assert out.returncode == 0 and lines, out.stdout + out.stderr
    res = lines[-1]
    assert res["relerr_logp_vs_oracle"] <= 1e-12 and res["relerr_grad_vs_oracle"] <= 1e-12
    assert res["bad_chain_logp"] == float("-inf") and res["bad_chain_grad_zero"]
