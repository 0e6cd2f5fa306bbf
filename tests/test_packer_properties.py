"""Property tests (hypothesis, CPU only) of the host packing entry points the Julia packer / the dataset would bind:
run offsets, user-block permutation, concat + pad, windows, shard-by-rank.  Oracle: oracle/cloze_oracle.py (numpy
restatement of transformer.py:53-98 and transformer.jl:144-163)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import cloze_oracle as oc
from recommendersystem_b200 import lib as L, packer


@pytest.fixture(scope="module", autouse=True)
def built():
    import os
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__ as g
        g.build()


# runs of user ids: ids may recur non-contiguously, runs may be length 1, trailing zero padding as the packer writes it
runs = st.lists(st.tuples(st.integers(0, 6), st.integers(1, 9)), min_size=0, max_size=40)


def expand(rs):
    return np.array([u for u, n in rs for _ in range(n)], dtype=np.int32)


@settings(max_examples=200, deadline=None)
@given(runs)
def test_offsets_are_the_run_boundaries(rs):
    u = expand(rs)
    off = packer.jagged_offsets(u)
    assert np.array_equal(off, oc.jagged_offsets(u))
    assert off[0] == 0 and off[-1] == len(u) and np.all(np.diff(off) > 0) or len(u) == 0
    for a, b in zip(off[:-1], off[1:]):          # constant inside a run, different across a boundary
        assert np.all(u[a:b] == u[a])
    for b in off[1:-1]:
        assert u[b - 1] != u[b]


@settings(max_examples=200, deadline=None)
@given(runs, st.randoms(use_true_random=False))
def test_block_shuffle_moves_whole_runs_in_order(rs, rnd):
    u = expand(rs)
    off = packer.jagged_offsets(u)
    order = list(range(len(off) - 1))
    rnd.shuffle(order)
    perm = packer.index_permutation(u, permutation=lambda n: np.asarray(order[:n], dtype=np.int64))
    assert np.array_equal(perm, oc.block_permutation(off, np.asarray(order, dtype=np.int64)))
    assert sorted(perm.tolist()) == list(range(len(u)))              # a permutation of the positions
    pos = 0
    for blk in order:                                                   # block by block, each run contiguous and in order
        a, b = int(off[blk]), int(off[blk + 1])
        assert perm[pos: pos + (b - a)].tolist() == list(range(a, b))
        pos += b - a
    d = {"userid": u.copy(), "x": np.arange(len(u), dtype=np.float32)}
    packer.block_shuffle(d, permutation=lambda n: np.asarray(order[:n], dtype=np.int64))
    assert np.array_equal(d["userid"], u[perm]) and np.array_equal(d["x"], perm.astype(np.float32))


@settings(max_examples=100, deadline=None)
@given(st.lists(st.integers(0, 17), min_size=1, max_size=12), st.sampled_from([1, 8, 64]))
def test_concat_pads_with_zeros_to_the_batch_multiple(lens, batch):
    rng = np.random.default_rng(sum(lens) + batch)
    users = [{"userid": np.full(n, i + 1, np.int32), "time": rng.standard_normal(n), "w": rng.standard_normal(n).astype(np.float32)}
             for i, n in enumerate(lens)]
    got = packer.concat(users, batch)
    ref = oc.concat_pad(users, batch)
    n = sum(lens)
    for k in ("userid", "time", "w"):
        assert got[k].dtype == users[0][k].dtype and len(got[k]) % batch == 0 and len(got[k]) - n < batch
        assert np.array_equal(got[k], ref[k])
        assert np.all(got[k][n:] == 0)
    if n:
        ws = list(packer.windows(got, batch))
        assert len(ws) == len(got["userid"]) // batch
        assert np.array_equal(np.concatenate([w["userid"] for w in ws]), got["userid"])


@given(st.integers(1, 8), st.integers(1, 4))
def test_shards_by_rank_partition_the_directories(world, per_rank):
    shards = [f"shard_{i:02d}" for i in range(world * per_rank)]
    seen = []
    for r in range(world):
        mine = packer.shards_for_rank(shards[::-1], r, world)      # order of discovery must not matter
        assert len(mine) == per_rank and all(int(s[-2:]) % world == r for s in mine)
        seen += mine
    assert sorted(seen) == shards
