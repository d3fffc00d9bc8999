"""Integer / layout work restated from the reference's own Java (bit-exact), SURVEY.md 3.3 and 8c(2)."""
import numpy as np
import pytest


def test_inner_neurons_matches_reference_java(oracle):
    # builder/function/MLP.java:28-42 evaluated by hand
    assert oracle.inner_neurons(10, 10, 0.65, 1) == [7]       # round(6.5)=7 (half-up)
    assert oracle.inner_neurons(8, 5, 0.65, 1) == [5]         # round(5.2)=5
    assert oracle.inner_neurons(2, 2, 0.65, 1) == [2]         # max(2, round(1.3))
    assert oracle.inner_neurons(10, 10, 0.65, 0) == []
    # nIL=2: i=0: 10 + (7-10)/(2)*1 = 10 + (-1) = 9 (Java int division truncates -1.5 -> -1); i=1: 7 + (10-7)/2*0 = 7
    assert oracle.inner_neurons(10, 10, 0.65, 2) == [9, 7]
    # nIL=3: nIL/2=1: i=0: 10 + (-3)/2*1 = 9 ; i=1: 7 + 3/2*0 = 7 ; i=2: 7 + 3/2*1 = 8
    assert oracle.inner_neurons(10, 10, 0.65, 3) == [9, 7, 8]
    assert oracle.inner_neurons(44, 11, 0.65, 1) == [29]      # round(28.6)


def test_genome_lengths_of_baseline_configs(oracle):
    assert oracle.config_model(1).genome_len() == 157   # 10-7-10: 11*7 + 8*10
    assert oracle.config_model(3).genome_len() == 16
    assert oracle.config_model(4).genome_len() == 75    # 8-5-5: 9*5 + 6*5
    assert oracle.count_weights([10, 7, 10]) == 157 and oracle.count_weights([8, 5, 5]) == 75


@pytest.mark.parametrize("shape,nv,ns,nr,nw", [("biped-4x3", 10, 180, 40, 24), ("worm-8x2", 16, 288, 64, 44),
                                                ("comb-7x2", 11, 198, 44, 20), ("box-1x1", 1, 18, 4, 0)])
def test_shape_counts(oracle, shape, nv, ns, nr, nw):
    import ctypes as C
    mask = np.ascontiguousarray(oracle.build_shape(shape))
    p = oracle.default_params()
    out = [C.c_int() for _ in range(4)]
    oracle.lib().vo_counts(C.byref(p), mask.shape[1], mask.shape[0], mask.ctypes.data_as(C.c_void_p),
                           *[C.byref(o) for o in out])
    assert [o.value for o in out] == [nv, ns, nr, nw]


def test_shapes(oracle):
    assert oracle.build_shape("biped-4x3").tolist() == [[1, 0, 0, 1], [1, 1, 1, 1], [1, 1, 1, 1]]
    assert oracle.build_shape("comb-7x2").tolist() == [[1, 0, 1, 0, 1, 0, 1], [1] * 7]
    assert oracle.build_shape("worm-8x2").sum() == 16


def test_episode_step_counts(oracle):
    # emulated `while (t < finalT) t += 1/60` in IEEE double
    assert oracle.episode_steps(30.0, 1.0) == (1801, 60, 1800)
    assert oracle.episode_steps(10.0, 1.0) == (600, 60, 599)
    assert oracle.episode_steps(60.0, 1.0) == (3601, 60, 3600)
    m = oracle.config_model(1, final_t=10.0)
    r, *_ = m.run(np.zeros(157))
    assert (r.n_steps, r.k_first, r.k_last) == (600, 60, 599)


def test_terrains(oracle):
    xs, ys = oracle.create_terrain("flat")
    assert xs.tolist() == [0, 10, 1990, 2000] and ys.tolist() == [100, 5, 5, 100]
    xs, ys = oracle.create_terrain("hilly-1-10-0")
    assert len(xs) == len(ys) and np.all(np.diff(xs) >= 1.0 - 1e-9) and xs[1] == 10 and ys[1] == 0 and ys[-1] == 100
    assert xs[2] == 10 + max(1.0, (0.8025330637390305 * 0.25 + 1) * 10)  # first Gaussian of Random(0)
    xs, ys = oracle.create_terrain("steppy-0.25-1-0")
    d = np.diff(xs)[1:-1]
    assert np.allclose(d[1::2], 0.5) and np.all(d[0::2] >= 1.0 - 1e-9)
    xs, ys = oracle.create_terrain("downhill-30")
    assert abs((ys[1] - ys[2]) - 1980 * 0.5) < 1e-9
