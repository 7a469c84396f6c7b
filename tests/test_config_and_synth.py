import numpy as np
import pytest

from limap_b200.config import DEFAULT_YAML_TRIANGULATION, make_tri_config
from limap_b200.synth import make_scene


def test_cpp_defaults_when_keys_missing():
    # ASSIGN_PYDICT_ITEM semantics (internal/helpers.h:25-27): missing keys keep the C++ defaults
    c = make_tri_config({})
    assert c.min_length_2d == 20.0 and c.line_tri_angle_threshold == 5.0 and c.min_num_outer_edges == 1
    assert c.linker2d.th_angle == 8.0 and c.linker2d.th_perp == 5.0 and c.linker2d.use_innerseg == 0
    assert c.linker3d.th_scaleinv == 0.01 and c.linker3d.use_innerseg == 1 and c.linker3d.use_perp == 0
    c = make_tri_config(DEFAULT_YAML_TRIANGULATION)
    assert c.min_length_2d == 0.0 and c.line_tri_angle_threshold == 1.0 and c.min_num_outer_edges == 0
    assert c.linker2d.th_angle == 5.0 and c.linker2d.th_overlap == 0.05 and c.linker2d.th_smartoverlap == 0.2
    assert c.linker3d.th_scaleinv == 0.015 and c.linker3d.th_smartangle == 2.0
    # unknown keys are ignored
    make_tri_config({"remerging": {"disable": False}, "filtering2d": {}, "not_a_key": 3})
    with pytest.raises(RuntimeError):
        make_tri_config({"merging_strategy": "nope"})


def test_scene_is_seed_deterministic_and_well_formed():
    a = make_scene(V=6, L=50, N=3, K=4, seed=5)
    b = make_scene(V=6, L=50, N=3, K=4, seed=5)
    assert np.array_equal(a.segs, b.segs) and np.array_equal(a.qvec, b.qvec)
    assert a.line_off[-1] == 6 * 50 and a.n_rows() == 6 * 3 * 50 * 4
    for i, m in a.matches.items():
        assert sorted(m.keys()) == sorted(a.neighbors[i])
        for g, rows in m.items():
            assert rows.dtype == np.int32 and rows.shape[1] == 2
            assert rows[:, 0].max() < 50 and rows[:, 1].max() < 50
    ng, off, pairs = a.flat_matches(int(a.img_ids[0]))
    assert list(ng) == sorted(ng) and off[-1] == len(pairs)
    # sharded match generation equals the corresponding part of the full scene
    c = make_scene(V=6, L=50, N=3, K=4, seed=5, match_views=range(2, 4))
    assert sorted(c.matches.keys()) == [2, 3]
    for g in c.matches[2]:
        assert np.array_equal(c.matches[2][g], a.matches[2][g])
