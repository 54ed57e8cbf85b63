"""The drop-in claim, proven against the reference's own program: oracle/_ref/mithra_ref_gpu is the UNMODIFIED reference
main() (src/mithra.cpp), job parser, parameter classes and Solver -- its own initialize() and its own solve() loop
(src/solver.cpp:1212-1418) -- linked with integration/mithra_gpu_dropin.cpp in place of src/fdtd.cpp / src/fdtdSC.cpp
and with libmithra_gpu.so (oracle/Makefile target ref_gpu; INTEGRATION.md option B).  It runs tests/jobs/micro-dropin.job
(Gaussian-beam seed with TF/SF injection, static undulator, three screens, power sampling, 203 field steps) and must
write the files the unmodified reference wrote for the same job on the CPU (tests/golden/micro-dropin.npz, made by
tests/golden/make_golden_dropin.py): same names, same number of lines, power within 1e-8, screen records within 1e-9."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "mithra_ref_gpu")
GOLDEN = os.path.join(ROOT, "tests", "golden", "micro-dropin.npz")


def _numbers(raw):
    return [np.array([float(t) for t in line.split()]) for line in bytes(raw).decode().splitlines()]


def test_golden_of_the_dropin_job_is_the_references_output():
    g = np.load(GOLDEN)
    names = sorted(g.files)
    assert names == ["txt/power-sampling/power-micro-0.txt"] + ["txt/screens/profile-p0-screen%d.txt" % i for i in range(3)]
    rows = _numbers(g[names[0]])
    assert len(rows) == 203 and all(r.size == 2 for r in rows)
    assert max(r[1] for r in rows) > 0.0                      # the seed reaches the power plane within the run
    assert all(len(_numbers(g[n])) > 100 for n in names[1:])   # most of the bunch crosses every screen


@pytest.mark.gpu
def test_unmodified_reference_main_and_loop_over_the_library(tmp_path):
    if not os.path.exists(EXE):
        pytest.fail("oracle/_ref/mithra_ref_gpu is missing: `make -C oracle ref_gpu` where /root/reference exists "
                    "(it travels with the gpurun snapshot)")
    out = subprocess.run([EXE, os.path.join(ROOT, "tests", "jobs", "micro-dropin.job")], cwd=str(tmp_path),
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert out.returncode == 0, out.stdout.decode()[-2000:]
    g = np.load(GOLDEN)
    for key in sorted(g.files):
        rel = key[len("txt/"):]
        fn = tmp_path / rel
        assert fn.exists(), rel
        got, want = _numbers(open(fn, "rb").read()), _numbers(g[key])
        assert len(got) == len(want), rel
        if "power" in rel:
            G, W = np.array(got), np.array(want)
            np.testing.assert_allclose(G[:, 0], W[:, 0], rtol=1e-14, err_msg=rel)          # abscissae: the reference's own clocks
            np.testing.assert_allclose(G[:, 1], W[:, 1], rtol=1e-8, atol=1e-12 * np.abs(W[:, 1]).max(), err_msg=rel)
        else:
            # the reference writes the crossings of a step in list order; so does the library (upload index)
            G, W = np.array(got), np.array(want)
            np.testing.assert_allclose(G, W, rtol=1e-9, atol=1e-12, err_msg=rel)


GOLDEN_BUNCH = os.path.join(ROOT, "tests", "golden", "micro-dropin-bunch.npz")


def test_golden_of_the_bunch_output_job_is_the_references_output():
    g = np.load(GOLDEN_BUNCH)
    dirs = sorted({k.split("/")[1] for k in g.files})
    assert dirs == ["bunch-profile", "bunch-sampling", "bunch-visualization", "power-map", "power-sampling", "screens"]
    assert sum(k.startswith("txt/power-map/") for k in g.files) == 6
    assert len(_numbers(g["txt/bunch-sampling/bunch.txt"])) >= 10
    assert sum(k.endswith(".vtu") for k in g.files) == 3 and sum(k.startswith("txt/bunch-profile/") for k in g.files) == 5


@pytest.mark.gpu
def test_reference_bunch_writers_run_on_the_bunch_of_the_device(tmp_path):
    """tests/jobs/micro-dropin-bunch.job = micro-dropin + power-visualization, bunch-sampling, bunch-profile and bunch-visualization groups, through
    oracle/_ref/mithra_ref_gpu: bunchSample / bunchProfile / bunchVisualize are the reference's OWN code (solver.cpp:1582-1792)
    working on chargeVectorn_, which the stub refreshes from the device in the field steps where one of them is due
    (integration/mithra_gpu_dropin.cpp refreshBunch).  Every file the unmodified reference wrote on the CPU must come out:
    same names, same line counts, XML lines identical, numbers to the printed digits (the positions carry the libm
    difference of the push, 1e-9)."""
    if not os.path.exists(EXE):
        pytest.fail("oracle/_ref/mithra_ref_gpu is missing: `make -C oracle ref_gpu` where /root/reference exists")
    out = subprocess.run([EXE, os.path.join(ROOT, "tests", "jobs", "micro-dropin-bunch.job")], cwd=str(tmp_path),
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert out.returncode == 0, out.stdout.decode()[-2000:]
    g = np.load(GOLDEN_BUNCH)
    for key in sorted(g.files):
        rel = key[len("txt/"):]
        fn = tmp_path / rel
        assert fn.exists(), rel
        ref = bytes(g[key]).decode().splitlines()
        got = open(fn).read().splitlines()
        assert len(got) == len(ref), rel
        if rel.endswith(".pvtu"):
            assert got == ref, rel
        elif rel.startswith("power-map"):
            # Solver::powerVisualize: the map is accumulated on the device, the .vts is the stub's writer (radiation.cpp:393-447)
            _compare_numeric_file(got, ref, rel)
        elif rel.endswith(".vtu"):
            rows_g, rows_r = [], []
            for a, b in zip(got, ref):
                if b.startswith("<"):
                    assert a == b, rel
                    continue
                ta, tb = a.split(), b.split()
                assert len(ta) == len(tb), rel
                if len(tb) == 3:
                    rows_g.append([float(x) for x in ta]); rows_r.append([float(x) for x in tb])
                else:
                    assert ta == tb, rel                                  # connectivity, offsets, types
            G, R = np.array(rows_g), np.array(rows_r)
            np.testing.assert_allclose(G, R, rtol=2e-4, atol=2e-4 * np.abs(R).max(), err_msg=rel)
        elif rel.startswith("bunch-sampling"):
            G = np.array([[float(x) for x in l.split()] for l in got]); R = np.array([[float(x) for x in l.split()] for l in ref])
            assert G.shape == R.shape and R.shape[1] == 13, rel
            scale = np.abs(R).max(axis=0)
            assert np.all(np.abs(G - R) <= 2e-4 * np.abs(R) + 1e-6 * scale + 1e-7), rel
        elif rel.startswith("bunch-profile"):
            assert float(got[0]) == pytest.approx(float(ref[0]), rel=1e-14)
            G = np.array([[float(x) for x in l.split()] for l in got[1:]]); R = np.array([[float(x) for x in l.split()] for l in ref[1:]])
            np.testing.assert_allclose(G, R, rtol=1e-8, atol=1e-12, err_msg=rel)
        elif "power" in rel:
            G, W = np.array(_numbers("\n".join(got).encode())), np.array(_numbers("\n".join(ref).encode()))
            np.testing.assert_allclose(G[:, 0], W[:, 0], rtol=1e-14, err_msg=rel)
            np.testing.assert_allclose(G[:, 1], W[:, 1], rtol=1e-8, atol=1e-12 * np.abs(W[:, 1]).max(), err_msg=rel)
        else:
            G, W = np.array(_numbers("\n".join(got).encode())), np.array(_numbers("\n".join(ref).encode()))
            np.testing.assert_allclose(G, W, rtol=1e-9, atol=1e-12, err_msg=rel)


GOLDEN_FIELD = os.path.join(ROOT, "tests", "golden", "micro-dropin-field.npz")


def _compare_numeric_file(got, ref, rel, frac_cols=()):
    """Text files of numbers with XML lines in between: XML identical, numbers to the 5 digits the reference prints
    (|g - r| <= 2e-4 |r| + 2e-4 max|column|).  frac_cols: columns that only have to agree on 90 % of the rows."""
    assert len(got) == len(ref), rel
    rows = {}
    for a, b in zip(got, ref):
        if b.lstrip().startswith("<"):
            assert a == b, rel
            continue
        ta, tb = a.split(), b.split()
        assert len(ta) == len(tb), rel
        rows.setdefault(len(tb), ([], []))
        rows[len(tb)][0].append([float(x) for x in ta]); rows[len(tb)][1].append([float(x) for x in tb])
    for width, (g_, r_) in rows.items():
        G, R = np.array(g_), np.array(r_)
        scale = np.abs(R).max(axis=0)
        ok = np.abs(G - R) <= 2e-4 * np.abs(R) + 2e-4 * scale + 1e-300
        for c in range(width):
            if c in frac_cols:
                assert ok[:, c].mean() > 0.9, (rel, c, ok[:, c].mean())
            else:
                assert ok[:, c].all(), (rel, c, int((~ok[:, c]).sum()))


def _compare_field_tree(root, g):
    for key in sorted(g.files):
        rel = key[len("txt/"):]
        fn = os.path.join(str(root), rel)
        assert os.path.exists(fn), rel
        ref = bytes(g[key]).decode().splitlines()
        got = open(fn).read().splitlines()
        if rel.endswith(".pvts"):
            assert got == ref, rel
        elif rel.startswith("field-profile"):
            # x y z Ay Ey Bx: the reference prints en_ / bn_ WITHOUT evaluating them (fdtd.cpp:1563-1577) -- whatever a node's
            # last evaluation left there; fresh on both sides are the end planes and the nodes the writers of this step touch
            _compare_numeric_file(got, ref, rel, frac_cols=(4, 5))
        elif rel.startswith("screens"):
            G, W = np.array(_numbers("\n".join(got).encode())), np.array(_numbers("\n".join(ref).encode()))
            np.testing.assert_allclose(G, W, rtol=1e-9, atol=1e-12, err_msg=rel)
        elif "power" in rel:
            G, W = np.array(_numbers("\n".join(got).encode())), np.array(_numbers("\n".join(ref).encode()))
            np.testing.assert_allclose(G[:, 0], W[:, 0], rtol=1e-14, err_msg=rel)
            np.testing.assert_allclose(G[:, 1], W[:, 1], rtol=1e-8, atol=1e-12 * np.abs(W[:, 1]).max(), err_msg=rel)
        else:
            _compare_numeric_file(got, ref, rel)


def test_golden_of_the_field_output_job_is_the_references_output():
    g = np.load(GOLDEN_FIELD)
    dirs = sorted({k.split("/")[1] for k in g.files})
    assert dirs == ["field-profile", "field-sampling", "field-visualization", "power-sampling", "screens"]
    assert sum(k.endswith(".vts") for k in g.files) == 5 and sum(k.endswith(".pvts") for k in g.files) == 3
    assert len(bytes(g["txt/field-sampling/field-0.txt"]).decode().splitlines()) >= 15


@pytest.mark.gpu
def test_reference_field_writers_run_on_the_potentials_of_the_device(tmp_path):
    """tests/jobs/micro-dropin-field.job = micro-dropin + field-sampling (three points, nine fields), field-visualization (two
    in-plane groups and all-domain) and field-profile, through oracle/_ref/mithra_ref_gpu: FdTd::fieldSample / fieldVisualize* /
    fieldProfile and the lazy fieldEvaluate behind them are the reference's OWN code (fdtd.cpp:818-1594), linked unchanged; the
    stub copies A^{n+1} and A^n back from the device in the field steps where one of them is due
    (integration/mithra_gpu_dropin.cpp refreshFields).  Every file of the CPU run must come out: same names, same line counts,
    XML identical, numbers to the printed digits."""
    if not os.path.exists(EXE):
        pytest.fail("oracle/_ref/mithra_ref_gpu is missing: `make -C oracle ref_gpu` where /root/reference exists")
    out = subprocess.run([EXE, os.path.join(ROOT, "tests", "jobs", "micro-dropin-field.job")], cwd=str(tmp_path),
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert out.returncode == 0, out.stdout.decode()[-2000:]
    _compare_field_tree(tmp_path, np.load(GOLDEN_FIELD))

