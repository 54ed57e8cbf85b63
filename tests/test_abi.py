"""The C-ABI library: it loads, exports every symbol include/mithra_gpu.h declares, the ctypes mirror has the same
struct sizes as the C header, and without a CUDA device it fails loudly instead of falling back to the CPU."""
import ctypes as C
import os
import re
import subprocess
import tempfile

from mithra_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mithra_gpu.h")


def test_library_exports_every_declared_symbol():
    lib = abi.load()
    text = open(HEADER).read()
    declared = set(re.findall(r"\b(mithra_gpu_[a-z_0-9]+)\s*\(", text))
    assert declared, "no declarations found in the header"
    assert declared == set(abi.SYMBOLS), "abi.SYMBOLS and the header disagree: %s" % (declared ^ set(abi.SYMBOLS))
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_layout_matches_header():
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "mithra_gpu.h"
int main () {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(MithraSignal), sizeof(MithraBeam), sizeof(MithraUndulator),
         sizeof(MithraPower), sizeof(MithraScreens), sizeof(MithraGpuParams), sizeof(MithraGpuCounters),
         offsetof(MithraGpuParams, seed), offsetof(MithraGpuParams, max_particles));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "s"), os.path.join(d, "s.c")])
        got = [int(x) for x in subprocess.check_output([os.path.join(d, "s")]).split()]
    want = [C.sizeof(abi.Signal), C.sizeof(abi.Beam), C.sizeof(abi.Undulator), C.sizeof(abi.Power), C.sizeof(abi.Screens),
            C.sizeof(abi.Params), C.sizeof(abi.Counters), abi.Params.seed.offset, abi.Params.max_particles.offset]
    assert got == want


def test_abi_version():
    assert abi.load().mithra_gpu_abi_version() == abi.ABI_VERSION


def test_no_cpu_fallback_without_device():
    lib = abi.load()
    if lib.mithra_gpu_device_count() > 0:
        return
    from tests import helpers
    p, _, _ = helpers.params_for("micro-nsfd")
    h = C.c_void_p()
    rc = lib.mithra_gpu_create(C.byref(p), C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CPU path" in lib.mithra_gpu_last_error()
