"""The int8 tensor-core GLM path of the library (csrc/glm_i8.cu): the default for Normal GLM statements with 33-64 columns
and >= 32 chains.  (i) ragged rows / chains / columns and a NaN chain against the oracle and against the FP64 DMMA path it
replaces (tools/i8_parity_check.py, one process per path because the switch is read once); (ii) the DMMA kernel stays covered at
the configs[1] shape: tests/test_gpu_scale.py once more with SMC_GLM_I8=0.  The prototype kernel it derives from is covered by
test_zz_gpu_prototype.py."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,chains", [(5000, 70), (16384, 256), (1000, 32), (70000, 33)])
def test_int8_glm_path_matches_the_oracle(rows, chains):
    env = dict(os.environ)
    env.pop("SMC_GLM_I8", None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "i8_parity_check.py"), str(rows), str(chains)],
                         capture_output=True, text=True, timeout=900, env=env)
    lines = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    assert out.returncode == 0 and lines, out.stdout + out.stderr
    res = lines[-1]
    assert res["relerr_logp_vs_oracle"] <= 1e-12 and res["relerr_grad_vs_oracle"] <= 1e-12
    assert res["relerr_logp_vs_dmma"] <= 1e-12 and res["relerr_grad_vs_dmma"] <= 1e-12
    assert res["bad_chain_logp"] == float("-inf") and res["bad_chain_grad_zero"]


@pytest.mark.parametrize("flush", ["64", "3", "1"])
def test_int8_pipeline_under_stress(flush):
    """many tiles per CTA, the residual scale growing along the rows (the sticky per-chain scales have to grow in the middle of a
    CTA's row range: read-back, re-slice), a short read-back period, every launch checked for pipeline time-outs, repeated
    evaluations bit for bit (tools/i8_stress.py)"""
    env = dict(os.environ, SMC_GLM_I8_FLUSH=flush, SMC_GLM_I8_CHECK="1", SMC_NO_GRAPHS="1")
    env.pop("SMC_GLM_I8", None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "i8_stress.py"), "400000", "1024", "6"],
                         capture_output=True, text=True, timeout=600, env=env)
    lines = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    assert out.returncode == 0 and lines, out.stdout[-1500:] + out.stderr[-1500:]
    res = lines[-1]
    assert res["bit_identical_repeats"]
    assert res["relerr_logp"] <= 1e-12 and res["relerr_grad"] <= 1e-12


def test_fp64_dmma_kernel_stays_covered_at_configs1():
    env = dict(os.environ, SMC_GLM_I8="0")
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_scale.py"), "-m", "gpu", "-q", "-x"],
                         capture_output=True, text=True, timeout=1200, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-1000:]
