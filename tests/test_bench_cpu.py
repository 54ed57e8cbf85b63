"""bench.py on a machine without a GPU: the reference arm (`--impl reference`: the unmodified reference alone on the host cores,
the one place besides the cpu_baseline leg where bench.py executes oracle/_ref) prints the contract's line, and our arm refuses
to run without a device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_dump")


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/ref_dump is built where /root/reference exists")
def test_reference_arm_prints_the_contract_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "sc-weak",
                                   "--steps", "1", "--warmup", "3"], cwd=ROOT, timeout=600).decode()
    lines = [ln for ln in out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cell-updates/s" and d["unit"] == "cell-updates/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 3
    assert d["value"] > 1e6 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("fdtdSC") and d["gpu_launches"] == 0


def test_our_arm_has_no_cpu_path():
    from mithra_b200 import abi
    if abi.load().mithra_gpu_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "sc-weak", "--steps", "1", "--warmup", "3"], cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert r.returncode != 0
    assert b"no CUDA device" in r.stdout and not any(ln.startswith(b"{") for ln in r.stdout.splitlines())
