"""Slab decomposition and chain sharding on the GPU (one device is enough: several
slab contexts share it).  Decomposed runs must be bit-identical to the plain
single-context run because Philox counters are keyed on global site indices."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

J = 0.1


def _single(cm, shape, occ, T, mu, seed, n_passes, sample_period=1):
    lat = cm.IsingLatticeGPU(shape, J=J)
    lat.set_conditions(T, mu)
    lat.seed_philox(seed)
    lat.upload(occ)
    lat.run_passes(n_passes, cm.MODE_CHECKERBOARD, sample_period)
    return lat


@pytest.mark.parametrize("transport", ["nccl", "peer"])
@pytest.mark.parametrize("n_slabs", [1, 2, 4])
def test_slabs_on_one_gpu_match_single_context(n_slabs, transport):
    import torch

    import casmcode_monte_b200 as cm
    from casmcode_monte_b200.parallel import GpuSlabEngine, slab_columns

    shape = [128, 48]
    n0, n1 = shape
    T, mu, seed, n_passes = 2633.0, 0.02, 20261017, 6
    occ = np.random.default_rng(1).choice(np.array([-1, 1], dtype=np.int32), size=n0 * n1)
    ref = _single(cm, shape, occ, T, mu, seed, n_passes)

    engines = []
    for r in range(n_slabs):
        cb, nc = slab_columns(n1, n_slabs, r)
        e = GpuSlabEngine(shape, cb, nc, J, T, mu, seed)
        e.upload(occ[n0 * cb : n0 * (cb + nc)])
        engines.append(e)

    def exchange(colour):
        for r, e in enumerate(engines):
            lo, hi = engines[(r - 1) % n_slabs], engines[(r + 1) % n_slabs]
            lo.halo(colour, 1).copy_(e.boundary(colour, 0))
            hi.halo(colour, 0).copy_(e.boundary(colour, 1))

    exchange(0)
    exchange(1)
    torch.cuda.synchronize()
    if transport == "peer":
        for r, e in enumerate(engines):
            e.lat.slab_ipc_attach(0, peer=engines[(r - 1) % n_slabs].lat)
            e.lat.slab_ipc_attach(1, peer=engines[(r + 1) % n_slabs].lat)
    S_series, B_series = [], []
    for t in range(n_passes):
        for colour in (0, 1):
            for e in engines:
                e.half_sweep(colour, t, sample=(colour == 1))
            if transport == "nccl":
                exchange(colour)
        torch.cuda.synchronize()
    got = np.concatenate([e.download() for e in engines])
    assert np.array_equal(got, ref.download())
    # per-slab fused samples add up to the single-context series
    S = sum(e.lat.samples_sb()[0] + e.lat.n_sites for e in engines) - n0 * n1
    B = sum(e.lat.samples_sb()[1] for e in engines)
    Sr, Br = ref.samples_sb()
    assert np.array_equal(S, Sr) and np.array_equal(B, Br)
    assert sum(e.counters()[1] for e in engines) == ref.counters()[1]
    assert sum(e.observables()[1] for e in engines) == int(Br[-1])


def test_library_side_pass_loop_and_restarted_pass_counter():
    """cmg_slab_run_passes (the pass loop inside the library, halo exchange fused into the
    kernels) on a ring of one slab attached to itself: bit-identical to the plain context.
    Re-running the trajectory from pass 0 (cmg_set_pass_counter) must neither hang nor
    change the result: the neighbour flags count fused half-sweeps, not pass indices."""
    import torch

    import casmcode_monte_b200 as cm
    from casmcode_monte_b200.parallel import GpuSlabEngine

    shape = [128, 48]
    n0, n1 = shape
    T, mu, seed, n_passes = 2633.0, -0.02, 4711, 7
    occ = np.random.default_rng(3).choice(np.array([-1, 1], dtype=np.int32), size=n0 * n1)
    ref = _single(cm, shape, occ, T, mu, seed, n_passes, sample_period=2)
    e = GpuSlabEngine(shape, 0, n1, J, T, mu, seed)
    with pytest.raises(cm.CmgError):
        e.lat.slab_run_passes(1)  # neighbours not attached
    for attempt in range(2):
        e.upload(occ)
        for colour in (0, 1):
            e.halo(colour, 1).copy_(e.boundary(colour, 0))
            e.halo(colour, 0).copy_(e.boundary(colour, 1))
        torch.cuda.synchronize()
        if attempt == 0:
            e.lat.slab_ipc_attach(0, peer=e.lat)
            e.lat.slab_ipc_attach(1, peer=e.lat)
        e.lat.set_pass_counter(0)
        e.lat.reset_counters()
        e.lat.clear_samples()
        e.lat.slab_run_passes(n_passes, 2)
        e.sync()  # a timed-out neighbour wait would be reported here
        assert np.array_equal(e.download(), ref.download())
        S, B = e.lat.samples_sb()
        assert np.array_equal(S, ref.samples_sb()[0]) and np.array_equal(B, ref.samples_sb()[1])
        assert e.counters()[1] == ref.counters()[1]
    # timing aid: with the exchange off the sweep still runs (stale halos), no waits
    e.lat.slab_set_halo_exchange(False)
    e.lat.slab_run_passes(2, 0)
    e.sync()
    e.lat.close()


def test_two_slabs_on_two_streams_run_their_own_pass_loops():
    """Two slab contexts of one GPU on two streams, each running cmg_slab_run_passes on its
    own: the kernels order themselves through the neighbour flags alone."""
    import torch

    import casmcode_monte_b200 as cm
    from casmcode_monte_b200.parallel import GpuSlabEngine, slab_columns

    shape = [128, 48]
    n0, n1 = shape
    T, mu, seed, n_passes = 2500.0, 0.01, 99, 5
    occ = np.random.default_rng(4).choice(np.array([-1, 1], dtype=np.int32), size=n0 * n1)
    ref = _single(cm, shape, occ, T, mu, seed, n_passes)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    engines = []
    for r in range(2):
        cb, nc = slab_columns(n1, 2, r)
        e = GpuSlabEngine(shape, cb, nc, J, T, mu, seed, stream=streams[r].cuda_stream)
        e.upload(occ[n0 * cb : n0 * (cb + nc)])
        engines.append(e)
    torch.cuda.synchronize()
    for colour in (0, 1):
        for r, e in enumerate(engines):
            engines[(r - 1) % 2].halo(colour, 1).copy_(e.boundary(colour, 0))
            engines[(r + 1) % 2].halo(colour, 0).copy_(e.boundary(colour, 1))
    torch.cuda.synchronize()
    for r, e in enumerate(engines):
        e.lat.slab_ipc_attach(0, peer=engines[(r - 1) % 2].lat)
        e.lat.slab_ipc_attach(1, peer=engines[(r + 1) % 2].lat)
    for e in engines:
        e.lat.slab_run_passes(n_passes, 1)  # asynchronous: both loops are in flight together
    for e in engines:
        e.sync()
    assert np.array_equal(np.concatenate([e.download() for e in engines]), ref.download())
    B = sum(e.lat.samples_sb()[1] for e in engines)
    assert np.array_equal(B, ref.samples_sb()[1])
    for e in engines:
        e.lat.close()


def test_resident_ring_over_slabs_matches_single_context():
    """k_ring2d on slab contexts (variant "ring2d"): the slab stays in shared memory and the
    outer tiles trade their edge columns with the neighbour slab's outer tiles through the
    mailboxes (stores into the neighbour's memory inside the cooperative kernel, stamps that
    count the half-sweeps of the whole trajectory, no zeroing, no flags).  (a) one slab whose
    ring closes on itself through the "remote" path; (b) two slabs of one GPU, 64 tiles each,
    on two streams -- both cooperative kernels are resident together -- over several calls
    and launches (passes per launch 3), with an upload in between."""
    import torch

    import casmcode_monte_b200 as cm
    from casmcode_monte_b200.parallel import GpuSlabEngine, slab_columns

    shape = [1024, 512]
    n0, n1 = shape
    T, mu, seed = 2633.0, 0.013, 4242
    rng = np.random.default_rng(8)
    occ = rng.choice(np.array([-1, 1], dtype=np.int32), size=n0 * n1)
    occ2 = rng.choice(np.array([-1, 1], dtype=np.int32), size=n0 * n1)

    def single(o, passes):
        lat = cm.IsingLatticeGPU(shape, J=J)
        lat.set_conditions(T, mu)
        lat.seed_philox(seed)
        lat.upload(o)
        lat.run_passes(passes, cm.MODE_CHECKERBOARD, 2)
        return lat

    for n_slabs in (1, 2):
        streams = [torch.cuda.Stream() for _ in range(n_slabs)]
        engines = []
        for r in range(n_slabs):
            cb, nc = slab_columns(n1, n_slabs, r)
            e = GpuSlabEngine(shape, cb, nc, J, T, mu, seed, stream=streams[r].cuda_stream)
            e.lat.set_kernel_variant("ring2d:rp=3:rt=64")
            e.upload(occ[n0 * cb : n0 * (cb + nc)])
            engines.append(e)
        torch.cuda.synchronize()
        for r, e in enumerate(engines):
            e.lat.slab_ipc_attach(0, peer=engines[(r - 1) % n_slabs].lat)
            e.lat.slab_ipc_attach(1, peer=engines[(r + 1) % n_slabs].lat)
        for n_passes in (7, 4):
            for e in engines:
                e.lat.slab_run_passes(n_passes, 2)  # asynchronous: all rings are in flight together
        for e in engines:
            e.sync()
            assert e.lat.kernel_variant == "ring2d"
        ref = single(occ, 11)
        assert np.array_equal(np.concatenate([e.download() for e in engines]), ref.download())
        assert np.array_equal(sum(e.lat.samples_sb()[1] for e in engines), ref.samples_sb()[1])
        assert sum(e.lat.counters()[1] for e in engines) == ref.counters()[1]
        # a new state: nothing of the old run may be taken for an edge of the new one
        for r, e in enumerate(engines):
            cb, nc = slab_columns(n1, n_slabs, r)
            e.upload(occ2[n0 * cb : n0 * (cb + nc)])
            e.lat.set_pass_counter(0)
        torch.cuda.synchronize()
        for e in engines:
            e.lat.slab_run_passes(5, 0)
        for e in engines:
            e.sync()
        ref2 = single(occ2, 5)
        assert np.array_equal(np.concatenate([e.download() for e in engines]), ref2.download())
        # the run left the neighbours' halos current: the bond sums of the slabs add up, and the
        # streaming kernel (fused halo exchange, flags) carries on from there
        assert sum(e.observables()[1] for e in engines) == ref2.sample_now()[1]
        for e in engines:
            e.lat.set_kernel_variant("auto")
        for e in engines:
            e.lat.slab_run_passes(2, 0)
        for e in engines:
            e.sync()
            assert e.lat.kernel_variant == "bulk2d"
        ref2.run_passes(2, cm.MODE_CHECKERBOARD, 0)
        assert np.array_equal(np.concatenate([e.download() for e in engines]), ref2.download())
        for e in engines:
            e.lat.close()


def test_slab_creation_errors():
    import casmcode_monte_b200 as cm

    with pytest.raises(cm.CmgError):
        cm.IsingLatticeGPU([100, 64], slab=(0, 32))  # n0 % 32 != 0
    with pytest.raises(cm.CmgError):
        cm.IsingLatticeGPU([128, 64], slab=(1, 32))  # odd col_begin
    with pytest.raises(cm.CmgError):
        cm.IsingLatticeGPU([128, 64], slab=(48, 32))  # beyond the lattice
    lat = cm.IsingLatticeGPU([128, 64], slab=(0, 32), J=J)
    lat.set_conditions(2000.0, 0.0)
    with pytest.raises(cm.CmgError):
        lat.run_passes(1)  # slabs are stepped half-sweep by half-sweep


def test_chain_grid_sharding_is_invisible():
    """A (T, mu) grid split over 1, 2 or 3 'ranks' (contexts) gives identical chains."""
    from casmcode_monte_b200.parallel import run_chain_grid

    conds = [(T, mu) for T in (1800.0, 2633.0, 3500.0) for mu in (-0.1, 0.0, 0.1)]
    shape = [64, 64]
    whole = run_chain_grid(conds, shape, n_passes=40, sample_period=2, seed=42)
    assert sorted(whole) == list(range(9)) and all(v["n_samples"] == 20 for v in whole.values())
    for world in (2, 3):
        merged = {}
        for rank in range(world):
            merged.update(run_chain_grid(conds, shape, n_passes=40, sample_period=2, seed=42, rank=rank, world_size=world))
        assert merged == whole
    # physics sanity: x increases with mu at fixed T, and is 1/2 at mu = 0 by symmetry within noise
    assert whole[3]["mean_param_composition"] < whole[5]["mean_param_composition"]


def test_full_size_65536_decomposition_is_invisible():
    """BASELINE config 5 at its full size (65536 x 65536 = 4.3e9 sites) on ONE GPU,
    with no host copy of the lattice: the undecomposed run and a two-slab run with
    fused peer halo pushes (same device) start from the same device-generated
    state (the draw is keyed on global site indices) and must produce the same
    integer sample series (S, B), acceptance counts and final observables."""
    import torch

    import casmcode_monte_b200 as cm
    from casmcode_monte_b200.parallel import GpuSlabEngine, slab_columns

    free, _ = torch.cuda.mem_get_info()
    if free < 40 << 30:
        pytest.skip("needs ~30 GB of device memory")
    shape = [65536, 65536]
    n0, n1 = shape
    n = n0 * n1
    T, mu, seed, n_passes = 2633.0, 0.01, 0xC0FFEE, 2

    lat = cm.IsingLatticeGPU(shape, J=J)
    lat.set_conditions(T, mu)
    lat.seed_philox(seed)
    lat.randomize(777, 0.5)
    S0, B0 = lat.sample_now()
    assert abs(S0) < 1e-3 * n and abs(B0) < 1e-3 * n  # an i.i.d. +-1 state
    lat.run_passes(n_passes, cm.MODE_CHECKERBOARD, 1)
    assert lat.kernel_variant == "bulk2d"
    S_ref, B_ref = lat.samples_sb()
    acc_ref = lat.counters()
    assert acc_ref[1] + acc_ref[2] == n_passes * n
    assert lat.sample_now() == (int(S_ref[-1]), int(B_ref[-1]))  # fused sampling == reduction of the state
    assert B_ref[-1] > B_ref[0] > B0  # relaxing towards order at T_c from a random state
    lat.close()
    del lat
    torch.cuda.empty_cache()

    engines = []
    for r in range(2):
        cb, nc = slab_columns(n1, 2, r)
        e = GpuSlabEngine(shape, cb, nc, J, T, mu, seed)
        e.lat.randomize(777, 0.5)
        engines.append(e)
    assert sum(e.observables()[0] + e.lat.n_sites for e in engines) - n == S0
    for colour in (0, 1):
        for r, e in enumerate(engines):
            engines[(r - 1) % 2].halo(colour, 1).copy_(e.boundary(colour, 0))
            engines[(r + 1) % 2].halo(colour, 0).copy_(e.boundary(colour, 1))
    torch.cuda.synchronize()
    for r, e in enumerate(engines):
        e.lat.slab_ipc_attach(0, peer=engines[(r - 1) % 2].lat)
        e.lat.slab_ipc_attach(1, peer=engines[(r + 1) % 2].lat)
    for t in range(n_passes):
        for colour in (0, 1):
            for e in engines:
                e.half_sweep(colour, t, sample=(colour == 1))
    torch.cuda.synchronize()
    S = sum(e.lat.samples_sb()[0] + e.lat.n_sites for e in engines) - n
    B = sum(e.lat.samples_sb()[1] for e in engines)
    assert np.array_equal(S, S_ref) and np.array_equal(B, B_ref)
    assert sum(e.counters()[1] for e in engines) == acc_ref[1]
    for e in engines:
        e.lat.close()
