"""Multi-GPU plumbing for the two ways this path shards:

1. independent chains -- the points of a (T, mu) grid are spread over ranks
   with no data-path communication (`shard_chains`, `run_chain_grid`);
2. slab decomposition -- one very large 2-d lattice is cut into column slabs,
   one per GPU, and after every coloured half-sweep each slab's two freshly
   updated boundary columns must reach its neighbours' halo columns
   (`SlabRing`).  Two transports:
     "nccl": torch.distributed send/recv of the boundary columns (device
             tensors over NVLink) -- the plumbing baseline;
     "peer": the half-sweep kernel itself stores its boundary results into the
             neighbour's halo (CUDA IPC / peer memory over NVLink) and raises a
             flag there; the next kernel spins on its own flag.  No host work
             and no collective between half-sweeps.
   Philox counters are keyed on global site indices, so the decomposed
   trajectory is bit-identical to the single-GPU one.

One process per GPU (torch.distributed).  torch is used for process groups,
streams and device tensors only; all lattice arithmetic is in the C-ABI library.
The ring logic is engine-agnostic so that it can be exercised with the gloo
backend on CPU (tests/ plug in a CPU engine of their own).
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------- chains
def shard_chains(n_chains: int, world_size: int, rank: int):
    """Contiguous block of chain indices owned by `rank` (sizes differ by <= 1)."""
    base, extra = divmod(n_chains, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return list(range(lo, hi))


def run_chain_grid(conditions, shape, n_passes, sample_period=1, J=0.1, seed=0xC0FFEE, device=0, rank=0, world_size=1, n_equil=0, dist=None):
    """Run the chains of `conditions` [(T, mu), ...] owned by this rank on one GPU
    and return {global chain index: dict(mean_x, mean_e_pot, n_accept, ...)}.
    With `dist` (an initialised torch.distributed) rank 0 gets the merged dict."""
    from . import MODE_CHECKERBOARD, Q_PARAM_COMPOSITION, Q_POTENTIAL_ENERGY, IsingLatticeGPU

    mine = shard_chains(len(conditions), world_size, rank)
    out = {}
    if mine:
        lat = IsingLatticeGPU(shape, n_chains=len(mine), device=device, J=J)
        for local, g in enumerate(mine):
            T, mu = conditions[g]
            lat.set_conditions(T, mu, chain=local)
        # every chain keeps the Philox stream of its GLOBAL index, so results do
        # not depend on how the grid is sharded
        lat.seed_philox(seed)
        lat.set_chain_offset(mine[0])
        if n_equil:
            lat.run_passes(n_equil, MODE_CHECKERBOARD, 0)
        lat.run_passes(n_passes, MODE_CHECKERBOARD, sample_period)
        for local, g in enumerate(mine):
            x = lat.samples(Q_PARAM_COMPOSITION, local)
            e = lat.samples(Q_POTENTIAL_ENERGY, local)
            n_pass, n_acc, n_rej = lat.counters(local)
            out[g] = {
                "T": conditions[g][0],
                "mu": conditions[g][1],
                "mean_param_composition": float(x.mean()) if x.size else float("nan"),
                "mean_potential_energy": float(e.mean()) if e.size else float("nan"),
                "n_samples": int(x.size),
                "n_accept": int(n_acc),
                "n_reject": int(n_rej),
                "checksum": int(lat.sample_now(local)[1]),
            }
        lat.close()
    if dist is not None and world_size > 1:
        gathered = [None] * world_size
        dist.all_gather_object(gathered, out)
        merged = {}
        for g in gathered:
            merged.update(g)
        return merged
    return out


# ----------------------------------------------------------------------------- slabs
def slab_columns(n1: int, world_size: int, rank: int):
    """(col_begin, n_cols) of this rank's slab; boundaries fall on even columns."""
    per = (n1 // 2) // world_size
    extra = (n1 // 2) % world_size
    lo = 2 * (rank * per + min(rank, extra))
    n = 2 * (per + (1 if rank < extra else 0))
    return lo, n


class GpuSlabEngine:
    """One column slab on one GPU (cmg_create_slab) behind the engine interface."""

    def __init__(self, global_shape, col_begin, n_cols, J, T, mu, seed, device=0, stream=None):
        import torch

        from . import IsingLatticeGPU

        self.torch = torch
        self.device = device
        self.lat = IsingLatticeGPU(global_shape, device=device, J=J, slab=(col_begin, n_cols))
        if stream is not None:
            self.lat.set_stream(stream)
        self.lat.set_conditions(T, mu)
        self.lat.seed_philox(seed)
        self.n0 = global_shape[0]
        self.n_cols = n_cols
        self._views = {}

    def _tensor(self, ptr, nbytes):
        key = (ptr, nbytes)
        if key not in self._views:

            class _Dev:
                pass

            d = _Dev()
            d.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
            self._views[key] = self.torch.as_tensor(d, device=f"cuda:{self.device}")
        return self._views[key]

    def upload(self, occ_local):
        self.lat.upload(occ_local)

    def download(self):
        return self.lat.download()

    def boundary(self, colour, side):
        return self._tensor(*self.lat.slab_boundary_ptr(colour, side))

    def halo(self, colour, side):
        return self._tensor(*self.lat.slab_halo_ptr(colour, side))

    def half_sweep(self, colour, pass_index, sample=False):
        self.lat.slab_half_sweep(colour, pass_index, sample)

    def run_passes(self, n_passes, sample_period=0):
        """Library-side pass loop (neighbours attached, halo exchange fused into the kernels)."""
        self.lat.slab_run_passes(n_passes, sample_period)

    def observables(self):
        """(ones-based S, B) partial sums of this slab for the current state."""
        return self.lat.sample_now()

    def counters(self):
        return self.lat.counters()

    def sync(self):
        self.lat.sync()


class SlabRing:
    """A ring of column slabs, one per rank, stepping a decomposed lattice.

    engine:    object with upload/download/boundary/halo/half_sweep (see GpuSlabEngine)
    dist:      torch.distributed module (initialised) or None for world_size 1
    transport: "nccl" (send/recv between half-sweeps; also what a gloo CPU test uses)
               or "peer" (kernel-side stores + flags; GPU engines only)
    """

    def __init__(self, engine, rank=0, world_size=1, dist=None, transport="nccl"):
        self.e = engine
        self.rank, self.world, self.dist = rank, world_size, dist
        self.transport = transport
        self.lo = (rank - 1) % world_size
        self.hi = (rank + 1) % world_size
        self.pass_index = 0

    # -- halo exchange of plane `colour`: my column 0 -> low neighbour's halo_hi,
    #    my last column -> high neighbour's halo_lo
    def exchange(self, colour):
        e = self.e
        if self.world == 1:
            e.halo(colour, 1).copy_(e.boundary(colour, 0))
            e.halo(colour, 0).copy_(e.boundary(colour, 1))
            return
        dist = self.dist
        ops = [
            dist.P2POp(dist.isend, e.boundary(colour, 0), self.lo),
            dist.P2POp(dist.isend, e.boundary(colour, 1), self.hi),
            dist.P2POp(dist.irecv, e.halo(colour, 0), self.lo),
            dist.P2POp(dist.irecv, e.halo(colour, 1), self.hi),
        ]
        if self.world == 2:
            # both neighbours are the same rank: order the pairs by tag-free convention
            # (send low first, receive the peer's "high" message into my low halo first)
            ops = [
                dist.P2POp(dist.isend, e.boundary(colour, 0), self.lo),
                dist.P2POp(dist.irecv, e.halo(colour, 1), self.hi),
                dist.P2POp(dist.isend, e.boundary(colour, 1), self.hi),
                dist.P2POp(dist.irecv, e.halo(colour, 0), self.lo),
            ]
        for r in dist.batch_isend_irecv(ops):
            r.wait()

    def attach_peers(self):
        """Exchange CUDA-IPC handles and let the kernels push their boundaries
        straight into the neighbours' halos (transport "peer")."""
        blob = self.e.lat.slab_ipc_export()
        if self.world == 1:
            self.e.lat.slab_ipc_attach(0, peer=self.e.lat)
            self.e.lat.slab_ipc_attach(1, peer=self.e.lat)
            return
        blobs = [None] * self.world
        self.dist.all_gather_object(blobs, blob)
        self.e.lat.slab_ipc_attach(0, handle=blobs[self.lo])
        self.e.lat.slab_ipc_attach(1, handle=blobs[self.hi])

    def prime(self):
        """Fill all halos from the neighbours' current boundaries (after upload)."""
        for colour in (0, 1):
            self.exchange(colour)
        if self.transport == "peer":
            if hasattr(self.e, "sync"):
                self.e.sync()
            if self.dist is not None and self.world > 1:
                self.dist.barrier()
            self.attach_peers()
            if self.dist is not None and self.world > 1:
                self.dist.barrier()

    def run_passes(self, n_passes, sample_period=0):
        if self.transport == "peer" and hasattr(self.e, "run_passes"):
            # the whole loop runs inside the C-ABI library: one call, no host work per half-sweep
            self.e.run_passes(n_passes, sample_period)
            self.pass_index += n_passes
            return
        for _ in range(n_passes):
            sample = sample_period > 0 and ((self.pass_index + 1) % sample_period) == 0
            for colour in (0, 1):
                self.e.half_sweep(colour, self.pass_index, sample and colour == 1)
                if self.transport != "peer":
                    self.exchange(colour)
            self.pass_index += 1

    def gather_lattice(self, n0, n1):
        """Full lattice (int32, column-major) on every rank -- for tests."""
        mine = np.ascontiguousarray(self.e.download(), dtype=np.int32)
        if self.world == 1:
            return mine
        parts = [None] * self.world
        self.dist.all_gather_object(parts, mine)
        return np.concatenate(parts)

    def global_observables(self):
        """(S, B) of the whole lattice: sum of the per-slab integer sums."""
        S, B = self.e.observables()
        if self.world == 1:
            return S, B
        parts = [None] * self.world
        self.dist.all_gather_object(parts, (int(S), int(B)))
        return sum(p[0] for p in parts), sum(p[1] for p in parts)
