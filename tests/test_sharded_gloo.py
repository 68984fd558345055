"""CPU, world_size 2 and 3 over gloo: the id-range sharding host logic of rii_b200/sharded.py (sample assembly, list
length exchange, per-shard IVF cut, all-gather + (distance, id) merge) reproduces the unsharded oracle."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_and_sample_ids():
    from rii_b200 import sharded
    assert sharded.shard_bounds(10, 4) == [0, 2, 5, 7, 10]
    ids = sharded.reference_sample_ids(5000, 20)   # min(N, 100*nlist) ids of the reference's shuffle
    assert ids.shape == (2000,) and len(set(ids.tolist())) == 2000 and ids.max() < 5000
    again = sharded.reference_sample_ids(5000, 20)
    assert np.array_equal(ids, again)
    assert sorted(sharded.reference_sample_ids(30, 20).tolist()) == list(range(30))


import pytest


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_shards_match_unsharded_oracle(world):
    port = 29500 + (os.getpid() + 7 * world) % 2000
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_shard_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (rank, o[-3000:])
        assert "rank %d ok" % rank in o
