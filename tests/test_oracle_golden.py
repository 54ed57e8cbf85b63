"""The CPU oracle (oracle/mithra_oracle.c) against the golden fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  This is what pins the oracle; the CUDA path is then checked against the oracle.

Potentials, currents and cell indices must be bit-identical (same association order, -ffp-contract=off, same libm);
so are particles, power and E/B, because oracle and reference run on the same host libm."""
import numpy as np
import pytest

from oracle import binding
from tests import helpers


@pytest.fixture(scope="module", params=helpers.JOBS)
def run(request):
    """Run the oracle for the 100 field steps of the fixture once per job, keeping the checkpoints."""
    job = request.param
    p, meta, g = helpers.params_for(job)
    o = binding.Oracle(p)
    helpers.start_from_golden(o, g)
    keep = {"job": job, "g": g, "p": p, "start": o.download_fields(("an", "anm1"))}
    for step in range(100):
        if step == 50:
            keep["p50"] = o.download_particles()
        if step == 99:
            o.fieldUpdate()
            names = ("anp1", "fnp1") if p.space_charge else ("anp1",)
            keep["ph_np1"] = o.download_fields(names)
            o.bunchUpdate()
            keep["ph_particles"] = o.download_particles()
            o.screenProfile()
            o.powerSample()
            o.powerVisualize()
            keep["ph_eb"] = o.download_eb()
            o.fieldShift()
            o.currentReset()
            o.currentUpdate()
            keep["ph_j"] = o.download_fields(names)
            o.advanceTime()
        else:
            helpers.solve_step(o)
    names = ("anp1", "an", "anm1") + (("fnp1", "fn", "fnm1") if p.space_charge else ())
    keep["end"] = o.download_fields(names)
    keep["p100"] = o.download_particles()
    keep["power"] = o.fetch_power()
    keep["pmap"] = o.fetch_power_map()
    keep["screens"] = [o.fetch_screen(s) for s in range(p.screens.N)] if p.screens.enabled else []
    o.close()
    return keep


def _check_array(g, key, step, arr, ncomp):
    a = arr.reshape(-1, ncomp)
    np.testing.assert_array_equal(a[g["idx"]], g["%s_s%s" % (key, step)], err_msg="%s step %s samples" % (key, step))
    np.testing.assert_array_equal(helpers.stats(a), g["%s_n%s" % (key, step)], err_msg="%s step %s sums" % (key, step))


def test_initial_fields(run):
    g = run["g"]
    _check_array(g, "an", 0, run["start"]["an"], 3)
    _check_array(g, "anm1", 0, run["start"]["anm1"], 3)


def test_particles_mid_run(run):
    np.testing.assert_array_equal(run["p50"], run["g"]["p50"])


def test_final_potentials_and_current(run):
    g, e = run["g"], run["end"]
    _check_array(g, "an", 100, e["an"], 3)
    _check_array(g, "anm1", 100, e["anm1"], 3)
    _check_array(g, "jn", 100, e["anp1"], 3)
    if run["p"].space_charge:
        _check_array(g, "fn", 100, e["fn"], 1)
        _check_array(g, "fnm1", 100, e["fnm1"], 1)
        _check_array(g, "rho", 100, e["fnp1"], 1)


def test_final_particles(run):
    np.testing.assert_array_equal(run["p100"], run["g"]["p100"])


def test_step99_phases(run):
    g = run["g"]
    _check_array(g, "ph99/anp1", "", run["ph_np1"]["anp1"], 3)
    _check_array(g, "ph99/jn", "", run["ph_j"]["anp1"], 3)
    if run["p"].space_charge:
        _check_array(g, "ph99/fnp1", "", run["ph_np1"]["fnp1"], 1)
        _check_array(g, "ph99/rho", "", run["ph_j"]["fnp1"], 1)
    np.testing.assert_array_equal(run["ph_particles"], g["ph99/particles"])
    en, bn, pic = run["ph_eb"]
    assert int(pic.sum()) == int(g["ph99/pic_count"][0])
    idx = g["ph99/pic_idx"]
    assert pic[idx].all()
    np.testing.assert_array_equal(en.reshape(-1, 3)[idx], g["ph99/en"])
    np.testing.assert_array_equal(bn.reshape(-1, 3)[idx], g["ph99/bn"])


def test_power_series(run):
    np.testing.assert_array_equal(run["power"], run["g"]["power"])


def test_power_map(run):
    """Solver::powerVisualize: the per-pixel map after 100 steps, bit for bit (jobs with a power-visualization group)."""
    if not run["p"].power_map.enabled:
        assert "pmap" not in run["g"].files
        return
    assert np.abs(run["g"]["pmap"]).max() > 0
    np.testing.assert_array_equal(run["pmap"], run["g"]["pmap"])


def test_screens(run):
    g = run["g"]
    for s, rec in enumerate(run["screens"]):
        ref = g["screen%d" % s]
        assert rec.shape == ref.shape
        # the reference writes 15 significant digits (solver.cpp:2183-2187)
        np.testing.assert_allclose(rec, ref, rtol=2e-15, atol=0)


def test_oracle_reproduces_the_files_of_the_reference_main_over_a_whole_job(tmp_path):
    """A whole job, not 100 steps of ref_dump: tests/jobs/micro-dropin.job (TF/SF seed, static undulator, three lab-frame screens,
    power sampling) as the UNMODIFIED reference main() ran it to its end on the CPU (tests/golden/micro-dropin.npz: the power file
    and the three screen files as written).  The oracle starts from the host's initialize() (bit-identical to the reference's on
    every fixture and shipped job, tests/test_host.py), marches the 203 field steps of Solver::solve and must give the numbers of
    those files to every printed digit: power rows (radiation.cpp:222-230) and screen records (solver.cpp:2229-2252)."""
    import ctypes as C
    import os
    import subprocess
    from mithra_b200 import abi, meta as mmeta
    root = helpers.ROOT
    exe = os.path.join(root, "mithra_b200", "host", "mithra_b200")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(exe)])
    pre = str(tmp_path / "h")
    subprocess.check_output([exe, os.path.join(root, "tests", "jobs", "micro-dropin.job"), "--dump-params", pre], cwd=str(tmp_path))
    rec = mmeta.read_records(pre + ".meta.bin")
    p = abi.Params.from_buffer_copy(rec["params0"].tobytes())
    p.max_particles = rec["particles"].size // 11 + 16
    p.max_screen_records = 1 << 16
    o = binding.Oracle(p)
    o.set_time(float(rec["time"][0]), float(rec["timeBunch"][0]), int(rec["nTime"][0]))
    o.upload_particles(rec["particles"].reshape(-1, 11))
    o.seedInitial()
    total, dt = float(rec["totalTime"][0]), float(rec["dt"][0])
    time, nsteps, tb = float(rec["time"][0]), 0, []
    while time < total:                                   # the loop condition of solver.cpp:1300 with the loop's own clock
        helpers.solve_step(o)
        time += dt
        nsteps += 1
    g = np.load(os.path.join(root, "tests", "golden", "micro-dropin.npz"))
    want = np.array([[float(x) for x in ln.split()] for ln in bytes(g["txt/power-sampling/power-micro-0.txt"]).decode().splitlines()])
    assert nsteps == want.shape[0] == 203
    power = np.asarray(o.fetch_power()).reshape(nsteps, -1)
    # the file prints 6 significant digits (default stream precision): compare at that resolution
    np.testing.assert_allclose(power[:, 0], want[:, 1], rtol=6e-6, atol=1e-300)
    assert want[:, 1].max() > 0.0
    for s in range(3):
        ref = np.array([[float(x) for x in ln.split()] for ln in bytes(g["txt/screens/profile-p0-screen%d.txt" % s]).decode().splitlines()])
        got = np.asarray(o.fetch_screen(s)).reshape(-1, 6)
        assert got.shape == ref.shape and ref.shape[0] > 100
        np.testing.assert_allclose(got, ref, rtol=6e-6, atol=1e-300)
    o.close()


def test_oracle_reproduces_the_power_maps_of_the_reference_main(tmp_path):
    """tests/jobs/micro-dropin-bunch.job through the unmodified reference main() (tests/golden/micro-dropin-bunch.npz): the six
    power-visualization files power-map/pmap-<nTime>.vts (radiation.cpp:324-450, written whenever fmod(time, rhythm) < dt).  The
    oracle marches the same job and its per-pixel map at those steps must be the printed numbers (5 significant digits), pixel by
    pixel in the file's order (j outermost)."""
    import os
    import subprocess
    from mithra_b200 import abi, meta as mmeta
    root = helpers.ROOT
    exe = os.path.join(root, "mithra_b200", "host", "mithra_b200")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(exe)])
    pre = str(tmp_path / "h")
    subprocess.check_output([exe, os.path.join(root, "tests", "jobs", "micro-dropin-bunch.job"), "--dump-params", pre], cwd=str(tmp_path))
    rec = mmeta.read_records(pre + ".meta.bin")
    p = abi.Params.from_buffer_copy(rec["params0"].tobytes())
    assert p.power_map.enabled == 1
    p.max_particles = rec["particles"].size // 11 + 16
    p.max_screen_records = 1 << 16
    g = np.load(os.path.join(root, "tests", "golden", "micro-dropin-bunch.npz"))
    due = sorted(int(k.split("pmap-")[1].split(".")[0]) for k in g.files if k.startswith("txt/power-map/"))
    assert due == [0, 40, 80, 120, 160, 200]
    o = binding.Oracle(p)
    o.set_time(float(rec["time"][0]), float(rec["timeBunch"][0]), int(rec["nTime"][0]))
    o.upload_particles(rec["particles"].reshape(-1, 11))
    o.seedInitial()
    N0, N1 = p.N0, p.N1
    seen = 0
    for step in range(due[-1] + 1):
        helpers.solve_step(o)                              # powerVisualize inside; the file of step n is written in step n
        if step in due:
            lines = bytes(g["txt/power-map/pmap-%d.vts" % step]).decode().splitlines()
            start = [i for i, ln in enumerate(lines) if "Name=\"power\"" in ln][0] + 1
            want = np.array([float(x) for x in lines[start:start + N0 * N1]]).reshape(N1, N0)        # j outermost, i inner
            got = np.asarray(o.fetch_power_map()).reshape(N0, N1).T
            np.testing.assert_allclose(got, want, rtol=6e-5, atol=1e-300, err_msg="pmap-%d" % step)
            seen += int(np.abs(want).max() > 0.0)
    assert seen >= 3                                       # the seed reaches the plane within the run
    o.close()
