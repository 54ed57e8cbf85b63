"""Cross-DEVICE correctness: one process per GPU, real cudaIpcOpenMemHandle mappings, ghost planes /
boundary currents / migrating particles as stores into peer memory over NVLink -- the slabs of 2 (and 4) GPUs must
reproduce the single-slab CPU oracle after 100 field steps, through the separate entry points and through
mithra_gpu_step (look-ahead, side streams).  Needs at least two visible GPUs; on a one-GPU box the tests are skipped
(tests/test_gpu_slabs.py runs the same exchange code with all slabs on one device)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    from mithra_b200 import abi
    return abi.load().mithra_gpu_device_count()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("job,world,fused", [("micro-nsfd", 2, False), ("micro-nsfd", 2, True), ("micro-sc", 2, True),
                                              ("micro-seeded", 4, True), ("micro-fd", 4, False)])
def test_slabs_on_different_devices_reproduce_the_single_slab_oracle(job, world, fused):
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "check_slabs_mp.py"), job] + (["--fused"] if fused else [])
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    text = r.stdout.decode()
    lines = [l for l in text.splitlines() if l.startswith("SLABS-MP ")]
    assert r.returncode == 0 and lines, text[-3000:]
    out = json.loads(lines[-1][len("SLABS-MP "):])
    assert out["ok"] and out["pids"] == world and len(set(out["devices"])) == world
    assert any(out["net_migration_per_slab"]) or job != "micro-nsfd"      # particles do change slabs in this job
