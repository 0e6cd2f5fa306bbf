"""Input pipeline on CPU: shard ownership by rank, whole-user block shuffle,
fixed-size windows - against a literal restatement of PretrainDataset (transformer.py:37-98)."""
import numpy as np

from recommendersystem_b200 import loader, packer, synthetic


def _users(n, rng):
    out, uid = [], 0
    for _ in range(n):
        m = int(rng.integers(3, 40))
        d = {}
        for k in synthetic.all_keys():
            dt = synthetic.key_dtype(k)
            d[k] = (rng.integers(0, 9, m) if np.issubdtype(dt, np.integer) else rng.random(m)).astype(dt)
        d["matchedid"] = np.arange(uid, uid + m, dtype=np.int32)   # a unique tag per token
        d["userid"] = np.ones(m, np.int32)
        uid += m
        out.append(d)
    return out, uid


def _reference_file_stream(d, perm_of_blocks, batch):
    """transformer.py:53-98 literally, with the block order injected."""
    arr = d["userid"]
    change = np.where(arr[:-1] != arr[1:])[0] + 1
    blocks = np.split(np.arange(len(arr)), change)
    p = np.concatenate([blocks[b] for b in perm_of_blocks])
    d = {k: v[p] for k, v in d.items()}
    return [{k: v[i:i + batch] for k, v in d.items()} for i in range(0, len(arr), batch)]


def test_loader_matches_reference_dataset(tmp_path):
    rng = np.random.default_rng(0)
    users, n_tok = _users(45, rng)
    loader.write_shards(str(tmp_path), users, n_shards=4, users_per_file=6, batch_size=64, seed=1)
    seen = []
    for rank in range(2):
        ld = loader.PretrainLoader(str(tmp_path), rank, 2, 64, seed=5, shuffle_files=False)
        assert len(ld.fns) == 4 and len({f.split("/")[-2] for f in ld.fns}) == 2                # two whole shard dirs
        assert {int(f.split("/")[-2]) % 2 for f in ld.fns} == {(rank + 1) % 2}                 # dir i -> rank i % world
        # literal reference on the same files with the same permutations
        ref_rng = np.random.default_rng(5)
        ref = []
        for fn in ld.fns:
            d = loader.read_shard_file(fn)
            n_blocks = len(packer.jagged_offsets(d["userid"])) - 1
            ref += _reference_file_stream(d, ref_rng.permutation(n_blocks), 64)
        got = list(ld)
        assert len(got) == len(ref) > 0
        for a, b in zip(got, ref):
            assert sorted(a) == sorted(b) and len(a) == 27
            for k in b:
                assert a[k].numpy().dtype == b[k].dtype and np.array_equal(a[k].numpy(), b[k]), k
            assert len(a["userid"]) == 64
        for a in got:
            real = a["userid"].numpy() > 0
            seen.append(a["matchedid"].numpy()[real])
    seen = np.sort(np.concatenate(seen))
    assert np.array_equal(seen, np.arange(n_tok))             # every token exactly once across the two ranks
