"""CPU-side checks of the measurement plumbing: the reference arm of bench.py runs the UNMODIFIED reference from oracle/_ref
(when the build recipe has produced it) and prints one JSON line with the contract's keys and the same `config` as the GPU arm;
oracle/_ref is verified against its manifest before it is trusted."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_tree_copy_is_verified_by_manifest(tmp_path, monkeypatch):
    from oracle import build_ref
    if not build_ref.available():
        pytest.skip("oracle/_ref not built here (no reference tree in this environment)")
    assert build_ref.verify()
    MatchModel, relax_matching, helper = build_ref.import_reference()
    assert MatchModel.__module__ == "dmm.modules.match_model" and callable(relax_matching)
    assert os.path.realpath(sys.modules["dmm.modules.match_model"].__file__).startswith(os.path.realpath(build_ref.OUT))
    # a tampered copy must be refused
    man = json.load(open(os.path.join(build_ref.OUT, "MANIFEST.json")))
    rel = "dmm/utils/match_helper.py"
    path = os.path.join(build_ref.OUT, rel)
    orig = open(path, "rb").read()
    try:
        open(path, "ab").write(b"\n# edited\n")
        assert not build_ref.verify()
    finally:
        open(path, "wb").write(orig)
    assert build_ref.verify() and rel in man["files"]


def test_reference_arm_prints_the_contract_line():
    import bench
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [x for x in r.stdout.splitlines() if x.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["config"] == bench.CONFIG
    assert d["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    from oracle import build_ref
    assert d["cpu_baseline"]["kind"] == ("reference" if build_ref.verify() else "port")
