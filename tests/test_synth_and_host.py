"""CPU: host-side logic — index-addressed generators shard consistently, the SoA layout helpers and the
stats combination used for the multi-GPU gather."""
import numpy as np

from cpprobotics_b200 import synth


def test_generators_are_index_addressed():
    n = 1000
    full = synth.ekf_inputs(n, n_steps=2)
    parts = [synth.ekf_inputs(250, i0=k * 250, n_steps=2) for k in range(4)]
    for a, *bs in zip(full, *parts):
        assert np.array_equal(a, np.concatenate(bs, axis=1))
    f = synth.pf_inputs(n, n_total=n)
    p = [synth.pf_inputs(500, i0=k * 500, n_total=n) for k in range(2)]
    assert np.array_equal(f[0], np.concatenate([p[0][0], p[1][0]], axis=1))
    assert np.array_equal(f[1], np.concatenate([p[0][1], p[1][1]]))
    course = synth.mpc_course()
    s, pi = synth.mpc_states(n, course=course)
    s2 = [synth.mpc_states(500, i0=k * 500, course=course) for k in range(2)]
    assert np.array_equal(s, np.concatenate([s2[0][0], s2[1][0]], axis=1))
    assert np.array_equal(pi, np.concatenate([s2[0][1], s2[1][1]]))


def test_ekf_inputs_are_well_conditioned():
    x, P, z, u = synth.ekf_inputs(2000)
    for i in range(0, 2000, 97):
        ev = np.linalg.eigvalsh(P[:, i].reshape(4, 4).T.astype(float))
        assert ev.min() > 0.05 and ev.max() < 5.0
    assert x.dtype == P.dtype == z.dtype == u.dtype == np.float32
    assert x.flags["C_CONTIGUOUS"] and P.shape == (16, 2000)


def test_different_seeds_differ_and_same_seed_repeats():
    a = synth.ekf_inputs(64, seed=1)[0]
    assert np.array_equal(a, synth.ekf_inputs(64, seed=1)[0])
    assert not np.array_equal(a, synth.ekf_inputs(64, seed=2)[0])


def test_engine_rejects_bad_arrays_before_touching_the_device():
    import pytest
    from cpprobotics_b200 import engine as E
    with pytest.raises(TypeError):
        E._ptr(np.zeros(4, np.float64), np.float32, device=False, name="x")
    with pytest.raises(ValueError):
        E._ptr(np.zeros((4, 8), np.float32)[:, ::2], np.float32, device=False, name="x")
    with pytest.raises(ValueError):
        E._ptr(np.zeros(4, np.float32), np.float32, device=True, name="x")
    with pytest.raises(ValueError):
        E._shape(np.zeros((3, 5), np.float32), 4, 5, "x")
