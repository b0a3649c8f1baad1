"""Multi-GPU slab decomposition check + timing (run under torchrun on N GPUs):
  torchrun --nproc-per-node N tools/slab_multi_gpu.py [n0 n1 n_passes]
Each rank owns a column slab of an n0 x n1 lattice.  Checks bit-identity of the
decomposed run against a single-GPU run of the same lattice (rank 0) for both
transports and reports attempts/s for each."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import casmcode_monte_b200 as cm
from casmcode_monte_b200.parallel import GpuSlabEngine, SlabRing, slab_columns


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    n1 = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    n_passes = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    check = n0 * n1 <= 8192 * 8192
    J, T, mu, seed = 0.1, 2633.0, 0.0, 0xC0FFEE
    cb, nc = slab_columns(n1, world, rank)
    rng = np.random.default_rng(1234)
    full = rng.choice(np.array([-1, 1], dtype=np.int32), size=n0 * n1) if check else None
    results = {}
    modes = [("nccl", "nccl", "auto"), ("peer", "peer", "auto"), ("ring", "peer", "ring2d")]
    if len(sys.argv) > 4:
        modes = [m for m in modes if m[0] in sys.argv[4].split(",")]
    for name, transport, variant in modes:
        stream = torch.cuda.current_stream()
        eng = GpuSlabEngine([n0, n1], cb, nc, J, T, mu, seed, device=local, stream=stream.cuda_stream)
        # "ring": the slab resident in shared memory (k_ring2d), its outer tiles trading edges
        # with the neighbour GPUs' outer tiles through peer memory -- one launch per block of passes
        eng.lat.set_kernel_variant(variant)
        if check:
            eng.upload(full[n0 * cb : n0 * (cb + nc)])
        else:
            eng.lat.randomize(99, 0.5)  # keyed on global site indices: one i.i.d. state over all slabs
        ring = SlabRing(eng, rank, world, dist, transport=transport)
        ring.prime()
        torch.cuda.synchronize()
        dist.barrier()
        ring.run_passes(3, sample_period=1)  # warm-up
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ring.run_passes(n_passes, sample_period=10)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        S, B = ring.global_observables()
        res = {"ms": ms, "attempts_per_s": float(n0) * n1 * n_passes / (ms * 1e-3), "S": int(S), "B": int(B)}
        if check:
            lattice = ring.gather_lattice(n0, n1)
            if rank == 0:
                ref = cm.IsingLatticeGPU([n0, n1], device=local, J=J)
                ref.set_conditions(T, mu)
                ref.seed_philox(seed)
                ref.set_kernel_variant("bulk2d")
                ref.upload(full)
                ref.run_passes(3, cm.MODE_CHECKERBOARD, 0)
                ref.run_passes(n_passes, cm.MODE_CHECKERBOARD, 0)
                res["bit_identical_to_single_gpu"] = bool(np.array_equal(lattice, ref.download()))
                res["B_single_gpu"] = int(ref.sample_now()[1])
                ref.close()
        res["kernel"] = eng.lat.kernel_variant
        results[name] = res
        eng.lat.close()
        dist.barrier()
    if rank == 0:
        print(json.dumps({"n_gpus": world, "lattice": [n0, n1], "n_passes": n_passes, "results": results}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
