"""include/vsr_host.hpp, the C++ host side above the C ABI (the reference's mapper / task interface: buildShape, buildSensorizingFunction,
getMapper(...).buildFor(target), Locomotion.apply -> Outcome; locomotion/Starter.java:402-436), driven by tests/cabi/host_mirror.cpp.
CPU: dimensions, genotype sizes and the Java error messages; no CPU fallback for the task itself.  GPU: the velocities of a population
mapped and evaluated through the C++ host side are bit-identical to the Python mirror's on the same genotypes."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    import __graft_entry__ as g
    from vsrevolution_b200 import _ffi
    if not os.path.exists(_ffi.LIB_PATH):
        g.build()
    exe = str(tmp_path / "host_mirror")
    libdir = os.path.dirname(_ffi.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cabi", "host_mirror.cpp"), "-L" + libdir, "-lvsr_eval",
                           "-Wl,-rpath," + libdir, "-o", exe])
    return exe


def _lcg_uniform(seed: int, count: int) -> np.ndarray:
    out, s = np.empty(count), seed & 0xFFFFFFFFFFFFFFFF
    for i in range(count):
        s = (s * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        out[i] = (s >> 11) * (1.0 / 9007199254740992.0)
    return out


def test_cpp_host_side_checks(tmp_path):
    exe = _build(tmp_path)
    import torch
    r = subprocess.run([exe, "check"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "host mirror checks ok" in r.stdout
    if torch.cuda.is_available():  # the task ran: one robot, then a generation behind the one-robot-per-call contract in one launch
        assert "gpu: velocity" in r.stdout and "lazy batching ok" in r.stdout and "average posture ok" in r.stdout
    else:
        assert "no gpu:" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("shape,sensors,mapper,terrain", [("biped-4x3", "uniform-a-0", "brainCentralized<mlp;r=0.65;nIL=1", "flat"),
                                                          ("worm-8x2", "", "brainPhaseVals;f=1", "hilly-1-10-0"),
                                                          ("comb-7x2", "uniform-t+vxy+a-0", "brainHomoDist;s=1<mlp;r=0.65;nIL=1", "steppy-0.25-1-0")])
def test_cpp_host_side_matches_the_python_mirror(tmp_path, shape, sensors, mapper, terrain):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from vsrevolution_b200.locomotion import population_evaluator
    exe = _build(tmp_path)
    out = str(tmp_path / "vel.bin")
    n, seed = 40, 2024
    r = subprocess.run([exe, "eval", shape, sensors, mapper, terrain, "3", "0.5", str(n), str(seed), out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    ev = population_evaluator(shape, sensors, mapper, terrain, 3.0, 0.5)
    assert f"genotype {ev.genome_len} robots {n}" in r.stdout
    g = (2.0 * _lcg_uniform(seed, n * ev.genome_len) - 1.0).reshape(n, ev.genome_len)
    assert np.array_equal(np.fromfile(out, dtype=np.float64), ev.evaluate(g))


@pytest.mark.gpu
def test_cpp_host_side_random_terrains_and_validation_match_the_python_mirror(tmp_path):
    """`-rnd` terrains drawn from the run's java.util.Random (Starter.java:89-93) and Starter.validate with a `broken-...` robot
    transformation (Starter.java:122-135), through the C++ host side and through the Python mirror: bit-identical velocities."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from vsrevolution_b200.locomotion import JavaRandomInts, build_locomotion_task, population_evaluator, validate
    exe = _build(tmp_path)
    out = str(tmp_path / "vel.bin")
    shape, sensors, mapper, n, seed = "biped-4x3", "uniform-a-0", "brainCentralized<mlp;r=0.65;nIL=1", 24, 77
    ev = population_evaluator(shape, sensors, mapper, "flat", 3.0, 0.5)
    g = (2.0 * _lcg_uniform(seed, n * ev.genome_len) - 1.0).reshape(n, ev.genome_len)
    robots = [ev._build(row) for row in g]
    # eval on hilly-1-10-rnd: the harness hands new Random(1) to buildLocomotionTask
    r = subprocess.run([exe, "eval", shape, sensors, mapper, "hilly-1-10-rnd", "3", "0.5", str(n), str(seed), out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    task = build_locomotion_task("hilly-1-10-rnd", 3.0, 0.5, random=JavaRandomInts(1))
    want = np.array([o.get_velocity() for o in task.apply_all(robots)])
    assert np.array_equal(np.fromfile(out, dtype=np.float64), want) and len(set(want)) > 1
    # validate: broken-0.4-rnd draws from new Random(5) of every validation, then the terrain seed
    for terrain, transformation in (("hilly-1-10-rnd", "broken-0.4-rnd"), ("flat", "broken-0.3-11"), ("downhill-20", "identity")):
        r = subprocess.run([exe, "validate", shape, sensors, mapper, terrain, "3", "0.5", str(n), str(seed), out, transformation, "5"],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        want = np.array([vo.outcome.get_velocity() for vo in validate(robots, terrain, transformation, 5, 3.0, 0.5)])
        assert np.array_equal(np.fromfile(out, dtype=np.float64), want), (terrain, transformation)
