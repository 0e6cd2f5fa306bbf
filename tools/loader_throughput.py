"""Host-side throughput of the input pipeline: how many interactions per second one loader thread
delivers as ready windows (file read + user-block shuffle through the C entry points + window slicing), against what one
B200 consumes (bench.py: ~0.83 M interactions/s).  Also times the reference's own Python permutation loop
(transformer.py:53-68, restated in oracle/cloze_oracle.py) on the same file.  CPU only:
    python tools/loader_throughput.py [positions_per_file]"""
import os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from recommendersystem_b200 import packer, synthetic
from recommendersystem_b200.loader import PretrainLoader

per_file = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * 1024 * 1024
cfg, rows = synthetic.make_config("C2")
S = cfg["max_sequence_length"]
batch = rows * S
smp = synthetic.HistorySampler(cfg, seed=3)
n_users = int(per_file / 190)
d, _ = smp.pack(n_users, batch)
with tempfile.TemporaryDirectory() as tmp:
    for s in (1, 2):
        os.makedirs(os.path.join(tmp, str(s)))
        np.savez(os.path.join(tmp, str(s), "1.npz"), **d)
    n = len(d["userid"])
    t0 = time.perf_counter()
    got = 0
    for w in PretrainLoader(tmp, 0, 1, batch, device=None, seed=0).host_batches():
        got += len(w["userid"])
    dt = time.perf_counter() - t0
    print(f"loader (1 background thread, npz files): {got} positions in {dt:.2f} s = {got / dt / 1e6:.2f} M interactions/s "
          f"({got / dt * 112 / 1e6:.0f} MB/s of packed arrays)")
    t0 = time.perf_counter(); p = packer.index_permutation(d["userid"], np.random.default_rng(0).permutation); t1 = time.perf_counter() - t0
    from oracle import cloze_oracle as oc
    t0 = time.perf_counter(); off = oc.jagged_offsets(d["userid"]); q = oc.block_permutation(off, np.random.default_rng(0).permutation(len(off) - 1)); t2 = time.perf_counter() - t0
    assert np.array_equal(p, q)
    print(f"user-block permutation of {n} positions ({len(off) - 1} users): C entry points {t1 * 1e3:.1f} ms, "
          f"the reference's Python loop (transformer.py:53-68) {t2 * 1e3:.1f} ms")
