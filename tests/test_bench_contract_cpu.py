"""bench.py's stdout contract, checked on the one workload that needs no GPU (the OpenAlex front
end): exactly ONE JSON line on stdout carrying the contract keys, for our arm and for the reference
arm; everything else (library banners, warnings) must go to stderr."""
import json
import os
import subprocess
import sys

from conftest import ROOT
from oracle import oa_jsonl as O


def _run(*extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "oa_jsonl", "--steps", "2", "--warmup", "1",
                        "--oa-records", "1500", *extra], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1, f"stdout must be exactly one JSON line, got {len(lines)}: {r.stdout[:300]!r}"
    return json.loads(lines[0])


def test_our_arm_prints_one_json_line_with_the_contract_keys():
    j = _run()
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "gpu_launches"):
        assert key in j, key
    assert j["value"] > 0 and j["unit"] == "MB/s" and j["dtype"] == "u8" and "workload" in j["config"]
    if O.reference_available():
        assert j["cpu_baseline"]["kind"] == "reference" and j["matches_reference_bytes"] is True


def test_reference_arm_prints_one_json_line():
    j = _run("--impl", "reference")
    assert j["impl"] == "reference"
    if O.reference_available():
        assert j["value"] > 0 and j["cpu_baseline"]["kind"] == "reference" and j["cpu_baseline"]["cores"] == 1
    else:
        assert "unavailable" in j
