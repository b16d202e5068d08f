"""bench.py end to end as a dry run (tests/dry_bench.py): the host code of the bench -- model construction, the training step,
graph-capture bookkeeping, per-kernel profile folding, roofline / e2e / clocks assembly -- runs through to ONE JSON line with
the keys the driver's contract names.  Values are not checked (no kernels run)."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.skipif(torch.cuda.is_available(), reason="dry runs are for GPU-less machines: with a device the "
                                "forwarded calls would launch kernels on host pointers")


@pytest.mark.parametrize("flags", [[], ["--dropout", "0.1", "--lengths", "librispeech"], ["--step", "fwd", "--precision", "tf32"],
                                   ["--workload", "conformer_m_ctc", "--no-graph", "--optimizer", "none"]])
def test_bench_main_prints_the_contract_line(flags):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dry_bench.py"), "--no-cpu-baseline", "--steps", "2",
                        "--warmup", "3"] + flags, capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and len(lines) == 1, r.stdout[-2000:] + r.stderr[-3000:]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert k in d, k
    assert d["metric"] == "speech_frames_per_sec" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3
    assert "workload" in d["config"] and "model" not in d["config"]
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"])
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["gpu_launches"] > 0
