"""The C ABI used from plain C (tests/cabi/eval_population.c): what a cgo / JNI / Panama binding does, with no Python in the process.
CPU: the harness compiles against include/vsr_eval.h alone, links libvsr_eval.so and, without a GPU, fails loudly with VSR_ECUDA
(no CPU fallback).  GPU: its fitness vector is bit-identical to the ctypes path on the same genomes."""
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
    exe = str(tmp_path / "eval_population")
    libdir = os.path.dirname(_ffi.LIB_PATH)
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cabi", "eval_population.c"), "-L" + libdir, "-lvsr_eval",
                           "-Wl,-rpath," + libdir, "-o", exe])
    return exe


def _lcg_uniform(seed: int, count: int) -> np.ndarray:
    out, s = np.empty(count), seed & 0xFFFFFFFFFFFFFFFF
    for i in range(count):
        s = (s * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        out[i] = (s >> 11) * (1.0 / 9007199254740992.0)
    return out


def test_c_harness_builds_and_has_no_cpu_fallback(tmp_path):
    exe = _build(tmp_path)
    import torch
    r = subprocess.run([exe, "biped-4x3", "uniform-a-0", "flat", "2", "0.5", "4", "7"], capture_output=True, text=True, timeout=300)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 1 and "vsr_ctx_create" in r.stderr and "-2" in r.stderr and "no CPU fallback" in r.stderr
    # bad names are rejected on the host side, before any device call
    r = subprocess.run([exe, "hexapod-4x3", "uniform-a-0", "flat", "2", "0.5", "4", "7"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "vsr_shape_mask" in r.stderr


@pytest.mark.gpu
def test_c_harness_matches_the_ctypes_path(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from vsrevolution_b200.locomotion import population_evaluator
    exe = _build(tmp_path)
    out = str(tmp_path / "fit.bin")
    n, seed = 50, 12345
    r = subprocess.run([exe, "biped-4x3", "uniform-a-0", "hilly-1-10-0", "3", "0.5", str(n), str(seed), out],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "mlp 10-7-10 genome 157" in r.stdout and "overflows 0" in r.stdout
    fit_c = np.fromfile(out, dtype=np.float64)
    ev = population_evaluator("biped-4x3", "uniform-a-0", "brainCentralized<mlp;r=0.65;nIL=1", "hilly-1-10-0", 3.0, 0.5)
    g = (2.0 * _lcg_uniform(seed, n * ev.genome_len) - 1.0).reshape(n, ev.genome_len)
    assert np.array_equal(fit_c, ev.evaluate(g))
