"""Property tests (hypothesis) of the host-side helpers around the hot path: shard partitioning, text formats."""
import os

import numpy as np
from hypothesis import given, settings, strategies as st

import limap.util.io as limapio
from limap_b200.dist import partition_by_cost, partition_views, slice_tracks


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(0, 10_000), min_size=1, max_size=200), st.integers(1, 9))
def test_partition_views_covers_in_order(weights, world):
    parts = partition_views(weights, world)
    assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == len(weights)
    assert all(a <= b for a, b in parts) and all(p[1] == q[0] for p, q in zip(parts[:-1], parts[1:]))


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(1, 500), min_size=0, max_size=300), st.integers(1, 9))
def test_partition_by_cost_is_a_balanced_permutation(costs, world):
    parts = partition_by_cost(costs, world)
    flat = np.concatenate(parts) if parts else np.zeros(0, np.int64)
    assert np.array_equal(np.sort(flat), np.arange(len(costs)))
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= 1
    if len(costs) >= world:
        tot = [sum(costs[i] for i in p) for p in parts]
        assert max(tot) - min(tot) <= max(costs)
    if costs:
        off = np.concatenate([[0], np.cumsum(costs)])
        new_off, (sub,) = slice_tracks(parts[0], off, np.arange(off[-1]))
        assert new_off[-1] == len(sub) == sum(costs[i] for i in parts[0])


@settings(max_examples=30, deadline=None)
@given(st.lists(st.lists(st.floats(-1e6, 1e6, allow_nan=False, width=64), min_size=4, max_size=4), min_size=1, max_size=20),
       st.integers(0, 10**6))
def test_segments_text_round_trip_is_exact(rows, img_id):
    import tempfile
    segs = np.array(rows, dtype=np.float64)
    with tempfile.TemporaryDirectory() as d:
        limapio.save_txt_segments(d, img_id, segs)
        assert os.listdir(d) == [f"segments_{img_id}.txt"]
        assert np.array_equal(limapio.read_txt_segments(d, img_id), segs)  # repr of a double parses back bit-exactly
