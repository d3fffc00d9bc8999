"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/vsr_eval.h
declares, its integer/naming layer is bit-identical to the oracle's tables, and it refuses to run without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def F():
    import __graft_entry__ as g
    from vsrevolution_b200 import _ffi
    if not os.path.exists(_ffi.LIB_PATH):
        g.build()
    return _ffi


def test_library_exports_every_declared_symbol(F):
    hdr = open(os.path.join(ROOT, "include", "vsr_eval.h")).read()
    declared = set(re.findall(r"\b(vsr_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = C.CDLL(F.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in vsr_eval.h but not exported"
    assert declared == set(F.EXPORTED_SYMBOLS)


def test_shapes_match_oracle_tables(F, oracle):
    for name in ("biped-4x3", "worm-8x2", "comb-7x2", "box-1x1", "box-3x5", "tripod-5x3", "biped-7x4", "comb-8x3"):
        assert np.array_equal(F.shape_mask(name), oracle.build_shape(name)), name
    with pytest.raises(F.VsrIllegalArgument):
        F.shape_mask("snake-3x3")


def test_terrains_bit_exact_vs_oracle(F, oracle):
    for name in ("flat", "hilly-1-10-0", "hilly-0.5-5-42", "steppy-0.25-1-0", "steppy-1-10-7", "downhill-30", "uphill-10"):
        xs, ys = F.terrain(name)
        oxs, oys = oracle.create_terrain(name)
        assert np.array_equal(xs, oxs) and np.array_equal(ys, oys), name
    with pytest.raises(F.VsrIllegalArgument):
        F.terrain("lunar-1")


def test_java_random_replay(F, oracle):
    r = oracle.JavaRandom(12345)
    exp = np.array([r.next_gaussian() for _ in range(64)])
    assert np.array_equal(F.java_random_gaussians(12345, 64), exp)
    assert F.java_random_gaussians(0, 1)[0] == 0.8025330637390305


def test_mlp_sizing_bit_exact(F, oracle):
    for n_in in range(1, 60, 3):
        for n_out in (1, 5, 10, 17):
            for r in (0.3, 0.65, 1.0, 1.7):
                for nil in (0, 1, 2, 3, 4, 5):
                    a = F.mlp_inner_neurons(n_in, n_out, r, nil)
                    assert a == oracle.inner_neurons(n_in, n_out, r, nil), (n_in, n_out, r, nil)
                    if all(v > 0 for v in a):
                        assert F.mlp_count_weights([n_in, *a, n_out]) == oracle.count_weights([n_in, *a, n_out])
    neurons = [10, 7, 10]
    # flat order: layer -> to-neuron -> (bias, from-neurons)
    assert F.mlp_weight_index(neurons, 0, 0, 0) == 0 and F.mlp_weight_index(neurons, 0, 1, 0) == 11
    assert F.mlp_weight_index(neurons, 1, 0, 0) == 77 and F.mlp_weight_index(neurons, 1, 9, 7) == 156


def test_episode_steps_bit_exact(F, oracle):
    for T, tr in ((30.0, 1.0), (10.0, 1.0), (60.0, 1.0), (5.0, 0.0), (2.0, 0.5), (0.5, 0.0)):
        assert F.episode_steps(T, tr) == oracle.episode_steps(T, tr)


def test_sensor_config(F, oracle):
    for name in ("uniform-a-0", "uniform-t+vxy+a-0", "uniform-a+vx+vy+r+t-0.01"):
        a, b = F.sensor_config(name), oracle.sensor_config(name)
        assert len(a) == len(b)
        for x, y in zip(a, b):
            for f in ("kind", "axes", "rotated", "soft_norm", "dom_min", "dom_max", "avg_window", "noise_sigma"):
                assert getattr(x, f) == getattr(y, f), (name, f)


def test_no_gpu_means_loud_failure_not_fallback(F):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    e = F.env_default()
    with pytest.raises(F.VsrError) as ei:
        F.Context(e, *F.terrain("flat"))
    assert ei.value.code == F.VSR_ECUDA and "no CPU fallback" in str(ei.value)


def test_create_rejects_bad_env(F):
    e = F.env_default()
    h = C.c_void_p()
    assert F.lib().vsr_ctx_create(C.byref(e), C.byref(h)) == F.VSR_EINVAL  # no terrain
    assert b"terrain" in F.lib().vsr_last_error(None)
    assert F.lib().vsr_ctx_create(None, C.byref(h)) == F.VSR_EINVAL
    assert F.lib().vsr_genome_len(None) == F.VSR_EINVAL


def test_jni_glue_type_checks_against_the_headers():
    """java/jni/vsr_eval_jni.c cannot be built here (no JDK), but it must at least type-check against include/*.h: compiled with
    -fsyntax-only against a minimal jni.h stand-in (tests/jni_stub/jni.h), and every native declared in VsrEval.java has its
    Java_..._<name> definition in the glue."""
    import subprocess
    src = os.path.join(ROOT, "java", "jni", "vsr_eval_jni.c")
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror=implicit-function-declaration", "-Werror=incompatible-pointer-types",
                        "-I" + os.path.join(ROOT, "tests", "jni_stub"), "-I" + os.path.join(ROOT, "include"), src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    java = open(os.path.join(ROOT, "java", "it", "units", "erallab", "gpu", "VsrEval.java")).read()
    natives = set(re.findall(r"public static native [\w\[\]]+ (\w+)\(", java))
    glue = set(re.findall(r"Java_it_units_erallab_gpu_VsrEval_(\w+)\(", open(src).read()))
    assert natives == glue and len(natives) >= 14


def test_sensor_config_voxels_matches_the_oracle_name_layer(F):
    """vsr_sensor_config_voxels (C++) against oracle/model.py sensor_config_voxels (Python): two independent restatements of
    RobotUtils.buildSensorizingFunction + PREDEFINED_SENSORS [U], compared field by field; unsupported vocabulary is loud."""
    import numpy as np
    from oracle import model as O
    for shape in ("biped-4x3", "worm-8x2", "comb-7x2", "box-1x1"):
        mask = O.build_shape(shape)
        for name in ("uniform-a-0", "uniform-t+vxy+a-0.01", "uniform-px+py+a-0", "uniform-l5+l1+lf5+cpg+r-0.02", "spinedTouch-f-f-0",
                     "spinedTouch-t-f-0.01", "spinedTouchSighted-t-f-0", "empty"):
            ls, off = F.sensor_config_voxels(name, mask)
            os_, ooff = O.sensor_config_voxels(name, mask)
            assert off.tolist() == ooff and len(ls) == len(os_), (shape, name)
            for a, b in zip(ls, os_):
                kb = (b.kind, b.axes, b.rotated, b.soft_norm, b.dom_min, b.dom_max, b.avg_window, b.noise_sigma, b.param, b.n_rays, b.side)
                ka = a.key()
                assert all((x == y) or (x != x and y != y) for x, y in zip(ka, kb)), (shape, name, ka, kb)  # NaN == NaN (px on a 1-wide grid)
    mask = O.build_shape("biped-4x3")
    for bad, code in (("uniform-m-0", F.VSR_EUNSUPPORTED), ("uniform-ax-0", F.VSR_EUNSUPPORTED), ("spinedTouch-t-t-0", F.VSR_EUNSUPPORTED),
                      ("spined-t-f-0", F.VSR_EINVAL), ("uniform-a", F.VSR_EINVAL)):
        with pytest.raises(F.VsrError) as ei:
            F.sensor_config_voxels(bad, mask)
        assert ei.value.code == code, (bad, ei.value.code)
    # the uniform entry point refuses the position-dependent sensors instead of returning one voxel's values for all
    with pytest.raises(F.VsrError):
        F.sensor_config("uniform-px+a-0")
