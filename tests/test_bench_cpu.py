"""bench.py without a GPU: the reference arm's JSON line on a shrunken workload, the byte model of the roofline, and
the refusal of the product arm to run without a CUDA device (there is no CPU fallback to time)."""
import argparse
import io
import json
import contextlib

import numpy as np
import pytest

import bench


def _args(**kw):
    d = dict(gpus=1, steps=1, warmup=0, impl="reference", no_cpu_baseline=False, no_full_fit=False, no_secondary=False)
    d.update(kw)
    return argparse.Namespace(**d)


def test_algorithmic_bytes_follow_survey_8d():
    assert bench.algorithmic_bytes_per_iter() == 200_000 * (4 * 128 + 18) == 106_000_000
    assert bench.algorithmic_bytes_per_iter(1_000_000, 32) == 146_000_000


def test_alpha_digest_is_a_function_of_the_values_only():
    a = np.linspace(0, 1, 1000)
    assert bench.alpha_digest(a) == bench.alpha_digest(a.copy()) == bench.alpha_digest(list(a))
    assert bench.alpha_digest(a) != bench.alpha_digest(np.nextafter(a, 2))
    assert bench.alpha_digest(a.astype(np.float32).astype(np.float64)[::-1][::-1]) == bench.alpha_digest(a.astype(np.float32))


def test_reference_arm_line(monkeypatch):
    """`--impl reference`: rank 0 prints one JSON line with the arm's own cpu_baseline / e2e blocks, the other ranks
    print nothing.  The workload is shrunk (the real one is n = 200 000) -- the code path is the same."""
    monkeypatch.setattr(bench, "N_ROWS", 600)
    monkeypatch.setattr(bench, "N_FEATURES", 8)
    monkeypatch.setattr(bench, "REF_ITERS", 5)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        bench.run_reference(_args(gpus=2), rank=1, world=2)
    assert out.getvalue() == ""
    with contextlib.redirect_stdout(out):
        bench.run_reference(_args(), rank=0, world=1)
    lines = [l for l in out.getvalue().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "smo_iterations_per_sec" and d["unit"] == "it/s"
    assert d["higher_is_better"] is True and d["scaling"] == "strong" and d["vs_baseline"] is None
    assert d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["iterations_per_step"] == 5 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == ("reference" if bench.load_reference() is not None else "port")
    assert cb["value"] == d["value"] and cb["cores"] >= 1 and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(SystemExit) as e:
        bench.run_ours(_args(impl="ours"), rank=0, world=1, local_rank=0)
    assert "no CPU fallback" in str(e.value)
