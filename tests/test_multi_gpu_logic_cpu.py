"""Multi-GPU host logic on CPU: chain sharding, slab geometry, and the slab ring
with a world_size-2 gloo process group (the GPU engines are swapped
for a CPU engine built on the oracle)."""
import json
import os
import socket
import subprocess
import sys

import pytest

from casmcode_monte_b200.parallel import shard_chains, slab_columns
from conftest import ROOT


def test_shard_chains_partition():
    for n, w in [(1024, 8), (10, 3), (3, 8), (128, 1)]:
        parts = [shard_chains(n, w, r) for r in range(w)]
        flat = [c for p in parts for c in p]
        assert flat == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert [len(shard_chains(1024, 8, r)) for r in range(8)] == [128] * 8  # BASELINE config 4


def test_slab_columns_cover_and_align():
    for n1, w in [(65536, 8), (65536, 2), (12, 2), (14, 3), (8192, 4)]:
        slabs = [slab_columns(n1, w, r) for r in range(w)]
        assert slabs[0][0] == 0 and sum(n for _, n in slabs) == n1
        for (b0, n0_), (b1, _) in zip(slabs, slabs[1:]):
            assert b0 + n0_ == b1
        assert all(b % 2 == 0 and n % 2 == 0 for b, n in slabs)  # local colour == global colour
    assert slab_columns(65536, 8, 3) == (3 * 8192, 8192)  # BASELINE config 5: 512 MiB int8 per GPU


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_slab_ring_gloo_matches_undecomposed_oracle(world):
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_slab_worker.py")], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[1][-2000:] for o in outs]
    line = [ln for ln in outs[0][0].splitlines() if ln.startswith("RESULT ")][0]
    r = json.loads(line[len("RESULT "):])
    assert r["identical"], r
    assert r["n_accept"] == r["n_accept_ref"]
    assert sum(n for _, n in r["slabs"]) == 12
