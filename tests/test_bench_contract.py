"""bench.py's contract on the CPU side: the reference arm (the C++ restatement on the host cores) prints one JSON line with the
keys the driver reads, the product arm refuses to run without a CUDA device (no CPU fallback), the roofline denominators come
from the committed measurements."""
import io
import json
import os
import sys
from contextlib import redirect_stdout

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_reference_arm_line(monkeypatch):
    monkeypatch.setattr(bench, "N_ROWS", 20_000)           # a bounded sample of the bounded sample: the contract, not the rate
    monkeypatch.delenv("RANK", raising=False)

    class A(object):
        gpus, steps, warmup = 1, 2, 1
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_reference(A)
    line = json.loads(buf.getvalue().strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["dtype"] == "f64"
    assert line["value"] > 0 and line["value"] == line["cpu_baseline"]["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "sample" in line["cpu_baseline"]
    assert "workload" in line["config"]


def test_reference_arm_runs_on_rank_zero_only(monkeypatch):
    monkeypatch.setenv("RANK", "1")

    class A(object):
        gpus, steps, warmup = 2, 1, 0
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_reference(A)
    assert buf.getvalue() == ""


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")

    class A(object):
        gpus, steps, warmup = 1, 1, 3
    with pytest.raises(SystemExit, match="no CUDA device"):
        bench._run_ours(A)


def test_roofline_denominators_come_from_committed_measurements():
    peak, src = bench.fp64_peak()
    assert 30.0 < peak < 45.0 and "fp64_peak_r01.json" in src
    probe = [json.loads(l) for l in open(os.path.join(ROOT, "profiles", "i8_umma_probe_r01.jsonl")) if l.startswith("{")]
    rates = [v for r in probe for k, v in r.items() if isinstance(v, (int, float)) and "pop" in k.lower()]
    assert rates and abs(max(rates) * (1000.0 if max(rates) < 100 else 1.0) - bench.I8_PEAK_TOPS) / bench.I8_PEAK_TOPS < 0.02, rates
