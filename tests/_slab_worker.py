"""Worker for the world_size-2 gloo test of the slab ring (CPU): each rank owns a
column slab held by a CPU engine built on the oracle's slab half-sweep, and the
ring logic under test (casmcode_monte_b200.parallel.SlabRing) moves the halos
with torch.distributed send/recv."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import _monte_oracle as orc  # noqa: E402

from casmcode_monte_b200.parallel import SlabRing, shard_chains, slab_columns  # noqa: E402


class OracleSlabEngine:
    """CPU stand-in for GpuSlabEngine (test infrastructure only)."""

    def __init__(self, n0, n1, col_begin, n_cols, J, T, mu, seed):
        self.n0, self.n1, self.col_begin, self.n_cols = n0, n1, col_begin, n_cols
        self.J, self.T, self.mu, self.seed = J, T, mu, seed
        self.occ = np.ones(n0 * n_cols, dtype=np.int32)
        h = n0 // 2
        # halo / boundary staging as plane-colour columns of h bytes, like the GPU engine
        self._halo = {(c, s): torch.zeros(h, dtype=torch.uint8) for c in (0, 1) for s in (0, 1)}
        self.halo_full = {0: np.ones(n0, dtype=np.int32), 1: np.ones(n0, dtype=np.int32)}
        self.n_accept = 0

    def upload(self, occ):
        self.occ = np.array(occ, dtype=np.int32)

    def download(self):
        return self.occ.copy()

    def _plane_column(self, col_values, j_global, colour):
        # bytes b=(1+s)/2 of the sites of `colour` in one column: i = 2p + ((j+colour)&1)
        par = (j_global + colour) & 1
        return ((col_values[par::2] + 1) // 2).astype(np.uint8)

    def boundary(self, colour, side):
        jl = 0 if side == 0 else self.n_cols - 1
        col = self.occ[self.n0 * jl : self.n0 * (jl + 1)]
        return torch.from_numpy(self._plane_column(col, self.col_begin + jl, colour).copy())

    def halo(self, colour, side):
        return self._halo[(colour, side)]

    def _sync_halos(self):
        for side, j in ((0, self.col_begin - 1), (1, self.col_begin + self.n_cols)):
            for colour in (0, 1):
                par = (j + colour) & 1
                self.halo_full[side][par::2] = 2 * self._halo[(colour, side)].numpy().astype(np.int32) - 1

    def half_sweep(self, colour, pass_index, sample=False):
        self._sync_halos()
        self.occ, acc = orc.checkerboard_half_sweep_slab(
            self.occ, self.halo_full[0], self.halo_full[1], self.n0, self.col_begin, self.n_cols, self.J, self.T, self.mu, self.seed, 0, pass_index, colour
        )
        self.n_accept += acc

    def observables(self):
        return 0, 0


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n0, n1, J, T, mu, seed, n_passes = 32, 12, 0.1, 2633.0, 0.03, 987654321, 4
    rng = np.random.default_rng(5)
    full = rng.choice(np.array([-1, 1], dtype=np.int32), size=n0 * n1)
    cb, nc = slab_columns(n1, world, rank)
    eng = OracleSlabEngine(n0, n1, cb, nc, J, T, mu, seed)
    eng.upload(full[n0 * cb : n0 * (cb + nc)])
    ring = SlabRing(eng, rank, world, dist, transport="nccl")
    ring.prime()
    ring.run_passes(n_passes)
    lattice = ring.gather_lattice(n0, n1)
    acc = [None] * world
    dist.all_gather_object(acc, int(eng.n_accept))
    if rank == 0:
        ref = orc.checkerboard_run([n0, n1], full, J, T, mu, seed, 0, 0, n_passes, 0)
        out = {
            "identical": bool(np.array_equal(lattice, ref["occupation"])),
            "n_accept": sum(acc),
            "n_accept_ref": int(ref["n_accept"]),
            "slabs": [slab_columns(n1, world, r) for r in range(world)],
            "chains": [shard_chains(10, world, r) for r in range(world)],
        }
        print("RESULT " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
