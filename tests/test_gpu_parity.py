"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU
oracle on the same seeded inputs.  Bit-exact for all integer / index work and
for every double that is a table lookup or a reference-order expression;
statistics within 1e-10 relative (reduction order differs)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

J = 0.1


@pytest.fixture(scope="module")
def cm():
    import casmcode_monte_b200 as m

    return m


def rand_occ(n, seed):
    rng = np.random.default_rng(seed)
    return rng.choice(np.array([-1, 1], dtype=np.int32), size=n)


def nsites(shape):
    return int(np.prod(shape))


# ---------------------------------------------------------------- tables ----
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("T,mu", [(2000.0, 0.0), (2633.0, 0.05), (300.0, -0.2), (1e5, 2.0), (5235.0, 1e-3)])
def test_tables_bit_exact(cm, oracle, dim, T, mu):
    shape = [4, 4] if dim == 2 else [4, 4, 4]
    lat = cm.IsingLatticeGPU(shape, J=J)
    lat.set_conditions(T, mu)
    dE, prob, thr = lat.tables()
    ref = oracle.accept_table(dim, J, T, mu)
    z = 2 * dim
    for b in range(2):
        for nu in range(z + 1):
            i = 2 * nu + b
            assert dE[i] == ref["dE"][b, nu]
            assert prob[i] == ref["prob"][b, nu]
            assert thr[i] == ref["thr_m1"][b, nu]


# ------------------------------------------------------ upload / download ----
@pytest.mark.parametrize("shape", [[2, 2], [6, 4], [25, 25], [7, 10], [64, 48], [4, 6, 8], [5, 4, 3], [32, 4, 6]])
def test_occupation_round_trip(cm, shape):
    n = nsites(shape)
    lat = cm.IsingLatticeGPU(shape, n_chains=2, J=J)
    assert np.array_equal(lat.download(0), np.ones(n, dtype=np.int32))  # fill_value=1 default
    a, b = rand_occ(n, 1), rand_occ(n, 2)
    lat.upload(a, 0)
    lat.upload(b, 1)
    assert np.array_equal(lat.download(0), a)
    assert np.array_equal(lat.download(1), b)
    lat.fill(-1, chain=0)
    assert np.array_equal(lat.download(0), -np.ones(n, dtype=np.int32))
    assert np.array_equal(lat.download(1), b)


def test_error_behaviour(cm):
    lat = cm.IsingLatticeGPU([6, 4], J=J)
    with pytest.raises(cm.CmgError):  # model.hh:56-58 size mismatch
        lat.upload(np.ones(23, dtype=np.int32))
    with pytest.raises(cm.CmgError):
        lat.upload(np.zeros(24, dtype=np.int32))  # values must be +-1
    with pytest.raises(cm.CmgError):
        lat.run_passes(1)  # conditions not set
    with pytest.raises(cm.CmgError):
        cm.IsingLatticeGPU([4, 4, 4, 4])  # model.hh:25-27
    with pytest.raises(cm.CmgError):
        lat.set_model(0.1, lattice_type=2)  # model.hh:175-177
    odd = cm.IsingLatticeGPU([25, 25], J=J)
    odd.set_conditions(2000.0, 0.0)
    with pytest.raises(cm.CmgError):
        odd.run_passes(1, cm.MODE_CHECKERBOARD)  # cannot two-colour an odd periodic lattice
    with pytest.raises(cm.CmgError):
        odd.run_passes(1, cm.MODE_SERIAL_REFERENCE)  # engine not seeded


# ------------------------------------------------------------------ probes ----
@pytest.mark.parametrize("shape", [[6, 4], [25, 25], [7, 10], [100, 100], [4, 6, 8], [5, 3, 7]])
@pytest.mark.parametrize("T,mu", [(2000.0, 0.0), (2633.0, 0.3)])
def test_delta_e_and_acceptance_bit_exact(cm, oracle, shape, T, mu):
    n = nsites(shape)
    occ = rand_occ(n, 5)
    lat = cm.IsingLatticeGPU(shape, J=J)
    lat.set_conditions(T, mu)
    lat.upload(occ)
    dE = lat.delta_e_probe()
    ref = oracle.potential_delta_all_sites(shape, occ, J, T, mu, True)
    assert np.array_equal(dE, ref)
    ref2 = oracle.potential_delta_all_sites(shape, occ, J, T, mu, False)
    assert np.array_equal(dE, ref2)
    u = np.random.default_rng(8).random(n)
    # put some uniforms exactly on / next to the thresholds
    _, prob, _ = lat.tables()
    u[: min(n, prob.size)] = np.clip(prob[: min(n, prob.size)], 0, np.nextafter(1.0, 0.0))
    acc = lat.accept_probe(u)
    assert np.array_equal(acc, oracle.accept_all_sites(ref, u, T))


def test_known_answers_through_the_device(cm, oracle):
    # tests/unit/monte/Ising_basic_semigrand_canonical_test.cpp:113-267
    lat = cm.IsingLatticeGPU([25, 25], J=J)
    lat.set_conditions(2000.0, 2.0)
    S, B = lat.sample_now()
    assert (S, B) == (625, 1250)
    x, ef, ep = oracle.observables_from_sums(S, B, 625, J, 2.0)
    assert x == 1.0 and math.isclose(ef, -2 * J) and math.isclose(ep, -2 * J - 2.0)
    dE = lat.delta_e_probe()
    assert np.all(dE == dE[0]) and math.isclose(dE[0], 8 * J + 2.0)


@pytest.mark.parametrize("shape", [[6, 4], [25, 25], [64, 48], [9, 14], [4, 6, 8], [5, 3, 7], [32, 8, 6]])
def test_integer_observables(cm, oracle, shape):
    n = nsites(shape)
    occ = rand_occ(n, 9)
    lat = cm.IsingLatticeGPU(shape, J=J)
    lat.upload(occ)
    assert lat.sample_now() == oracle.integer_observables(shape, occ)


def test_line_dots_nonlist_energy(cm, oracle):
    shape = [10, 12]
    occ = rand_occ(120, 4)
    lat = cm.IsingLatticeGPU(shape, J=J)
    lat.upload(occ)
    rows, cols = lat.line_dots()
    e = 0.0
    for d in rows:  # model.hh:273-285 order: rows then columns, -J * dot each
        e += -J * float(d)
    for d in cols:
        e += -J * float(d)
    assert e == oracle.formation_energy(shape, occ, J, False)[0]


# ------------------------------------------------- checkerboard trajectories ----
def run_cb(cm, shape, occ, T, mu, seed, n_passes, variant, n_chains=1, chain_conditions=None, sample_period=1):
    lat = cm.IsingLatticeGPU(shape, n_chains=n_chains, J=J)
    if chain_conditions is None:
        lat.set_conditions(T, mu)
    else:
        for ch, (t, m) in enumerate(chain_conditions):
            lat.set_conditions(t, m, chain=ch)
    lat.seed_philox(seed)
    lat.set_kernel_variant(variant)
    for ch in range(n_chains):
        lat.upload(occ[ch] if n_chains > 1 else occ, ch)
    lat.run_passes(n_passes, cm.MODE_CHECKERBOARD, sample_period)
    return lat


@pytest.mark.parametrize("shape", [[2, 2], [6, 4], [10, 14], [100, 100], [64, 48], [4, 6, 8], [2, 2, 2], [32, 4, 6]])
@pytest.mark.parametrize("T,mu", [(2633.0, 0.0), (1500.0, 0.1)])
def test_generic_kernel_matches_oracle(cm, oracle, shape, T, mu):
    n = nsites(shape)
    occ = rand_occ(n, 21)
    lat = run_cb(cm, shape, occ, T, mu, 0xC0FFEE, 6, "generic")
    ref = oracle.checkerboard_run(shape, occ, J, T, mu, 0xC0FFEE, 0, 0, 6, 1)
    assert np.array_equal(lat.download(), ref["occupation"])
    S, B = lat.samples_sb()
    assert np.array_equal(S, ref["S"]) and np.array_equal(B, ref["B"])
    n_pass, n_acc, n_rej = lat.counters()
    assert (n_pass, n_acc, n_rej) == (6, ref["n_accept"], ref["n_reject"])
    # sampled doubles follow the reference expression order bit-for-bit
    assert np.array_equal(lat.samples(cm.Q_PARAM_COMPOSITION), ref["param_composition"])
    assert np.array_equal(lat.samples(cm.Q_FORMATION_ENERGY), ref["formation_energy"])
    assert np.array_equal(lat.samples(cm.Q_POTENTIAL_ENERGY), ref["potential_energy"])


@pytest.mark.parametrize("shape", [[32, 2], [64, 48], [96, 34], [256, 256], [32, 6]])
@pytest.mark.parametrize("js", [0, 2, 5, 16])
def test_bulk2d_matches_oracle(cm, oracle, shape, js):
    n = nsites(shape)
    occ = rand_occ(n, 22)
    T, mu = 2633.0, 0.02
    variant = "bulk2d" if js == 0 else f"bulk2d:js={js}"
    lat = run_cb(cm, shape, occ, T, mu, 12345678901234567, 4, variant, sample_period=2)
    assert lat.kernel_variant == "bulk2d"
    ref = oracle.checkerboard_run(shape, occ, J, T, mu, 12345678901234567, 0, 0, 4, 2)
    assert np.array_equal(lat.download(), ref["occupation"])
    S, B = lat.samples_sb()
    assert np.array_equal(S, ref["S"]) and np.array_equal(B, ref["B"])
    assert lat.counters()[1] == ref["n_accept"]


@pytest.mark.parametrize(
    "shape,variant",
    [
        ([64, 48], "tile2d"),  # whole lattice in one tile: periodic columns, many passes per launch
        ([256, 256], "tile2d"),
        ([64, 6], "tile2d:nt=256"),
        ([4096, 64], "tile2d"),  # 64 one-column tiles with 4-column halos (2 passes per launch)
        ([4096, 64], "tile2d:p=1"),
        ([4096, 64], "tile2d:p=3:nt=1024"),
        ([1024, 512], "tile2d:p=2"),
        ([1024, 514], "tile2d:p=4:nt=256"),  # uneven tile widths
    ],
)
def test_tile2d_matches_oracle(cm, oracle, shape, variant):
    n = nsites(shape)
    occ = rand_occ(n, 24)
    T, mu = 2633.0, 0.02
    lat = run_cb(cm, shape, occ, T, mu, 424242, 5, variant, sample_period=2)
    assert lat.kernel_variant == "tile2d"
    ref = oracle.checkerboard_run(shape, occ, J, T, mu, 424242, 0, 0, 5, 2)
    assert np.array_equal(lat.download(), ref["occupation"])
    S, B = lat.samples_sb()
    assert np.array_equal(S, ref["S"]) and np.array_equal(B, ref["B"])
    assert lat.counters()[1] == ref["n_accept"]
    # continue with sampling every pass: the schedule phase carries over launches
    lat.run_passes(3, cm.MODE_CHECKERBOARD, 1)
    ref2 = oracle.checkerboard_run(shape, ref["occupation"], J, T, mu, 424242, 0, 5, 3, 1)
    assert np.array_equal(lat.download(), ref2["occupation"])
    S2, B2 = lat.samples_sb(first=len(S))
    assert np.array_equal(S2, ref2["S"]) and np.array_equal(B2, ref2["B"])


@pytest.mark.parametrize(
    "shape,variant",
    [
        ([1024, 128], "ring2d"),  # 4 tiles of 32 columns, 16 column groups of 2
        ([1024, 200], "ring2d:rp=2"),  # uneven tile widths, several launches
        ([2048, 96], "ring2d"),  # 8 column groups
        ([4096, 64], "ring2d:rp=1"),  # 4 column groups; one pass per launch
        ([4096, 300], "ring2d"),  # 37 tiles of 8-9 columns
    ],
)
def test_ring2d_matches_oracle(cm, oracle, shape, variant):
    # the lattice stays in shared memory across the launch; tile edges travel
    # through global memory under release/acquire flags
    n = nsites(shape)
    occ = rand_occ(n, 25)
    T, mu = 2633.0, 0.02
    lat = run_cb(cm, shape, occ, T, mu, 424242, 5, variant, sample_period=2)
    assert lat.kernel_variant == "ring2d"
    ref = oracle.checkerboard_run(shape, occ, J, T, mu, 424242, 0, 0, 5, 2)
    assert np.array_equal(lat.download(), ref["occupation"])
    S, B = lat.samples_sb()
    assert np.array_equal(S, ref["S"]) and np.array_equal(B, ref["B"])
    assert lat.counters()[1] == ref["n_accept"]
    lat.run_passes(3, cm.MODE_CHECKERBOARD, 1)
    ref2 = oracle.checkerboard_run(shape, ref["occupation"], J, T, mu, 424242, 0, 5, 3, 1)
    assert np.array_equal(lat.download(), ref2["occupation"])
    S2, B2 = lat.samples_sb(first=len(S))
    assert np.array_equal(S2, ref2["S"]) and np.array_equal(B2, ref2["B"])


def test_ring2d_two_chains_and_ties(cm, oracle):
    # two lattices share the cooperative grid (grid.y = chain, separate mailboxes);
    # 2 x 1024 x 512 x 8 passes = 8.4 M attempts -> a few hundred threshold ties
    shape = [1024, 512]
    n = nsites(shape)
    occ = [rand_occ(n, 41), rand_occ(n, 42)]
    conds = [(2633.0, 0.013), (2200.0, -0.02)]
    lat = run_cb(cm, shape, occ, None, None, 31337, 8, "ring2d", n_chains=2, chain_conditions=conds, sample_period=4)
    assert lat.kernel_variant == "ring2d"
    for ch, (T, mu) in enumerate(conds):
        ref = oracle.checkerboard_run(shape, occ[ch], J, T, mu, 31337, ch, 0, 8, 4)
        assert np.array_equal(lat.download(ch), ref["occupation"])
        S, B = lat.samples_sb(ch)
        assert np.array_equal(S, ref["S"]) and np.array_equal(B, ref["B"])
        assert lat.counters(ch)[1] == ref["n_accept"]


def test_ring2d_rejects_lattices_it_cannot_hold(cm):
    lat = cm.IsingLatticeGPU([96, 64], J=J)
    lat.set_conditions(2633.0, 0.0)
    lat.set_kernel_variant("ring2d")
    with pytest.raises(cm.CmgError, match="ring2d does not fit"):
        lat.run_passes(1, cm.MODE_CHECKERBOARD, 0)


def test_threshold_ties_take_the_exact_path(cm, oracle):
    # mu chosen so that a threshold's top half is hit often is impossible to arrange
    # (probability 2^-15 per site); instead run enough sites that ties occur:
    # 1024x1024 x 8 passes = 8.4M attempts -> ~256 expected ties; every one must
    # resolve exactly as the oracle's 32-bit compare does.
    shape = [1024, 1024]
    occ = rand_occ(nsites(shape), 25)
    for variant in ("tile2d", "bulk2d"):
        lat = run_cb(cm, shape, occ, 2633.0, 0.013, 31337, 8, variant, sample_period=8)
        ref = oracle.checkerboard_run(shape, occ, J, 2633.0, 0.013, 31337, 0, 0, 8, 8)
        assert np.array_equal(lat.download(), ref["occupation"])
        assert lat.counters()[1] == ref["n_accept"]


# the last three take the layer-paired warp mapping (several strips per warp, n2 % (2 * strips per warp) == 0)
@pytest.mark.parametrize("shape", [[32, 4, 2], [64, 6, 4], [32, 10, 8], [512, 8, 4], [256, 6, 8], [512, 22, 8]])
def test_bulk3d_matches_oracle(cm, oracle, shape):
    n = nsites(shape)
    occ = rand_occ(n, 23)
    T, mu = 5235.0, 0.05
    lat = run_cb(cm, shape, occ, T, mu, 99, 3, "auto")
    assert lat.kernel_variant == "bulk3d"
    ref = oracle.checkerboard_run(shape, occ, J, T, mu, 99, 0, 0, 3, 1)
    assert np.array_equal(lat.download(), ref["occupation"])
    S, B = lat.samples_sb()
    assert np.array_equal(S, ref["S"]) and np.array_equal(B, ref["B"])
    assert lat.counters()[1] == ref["n_accept"]


# k_halfsweep_tma3d: K = 8 (n0 = 512) or 4 (n0 = 1024) layers per CTA, columns by bulk copies.
# Shapes: one layer group that is its own k-neighbour (n2 = K), two groups, short strips
# (n1 = 2: one strip of two columns, the column after the last is column 0), n1 = 22 (ragged
# strips, start parities differ between the warps of a CTA), two chains.
@pytest.mark.parametrize("shape,chains", [([512, 2, 8], 1), ([512, 8, 8], 1), ([512, 22, 16], 1), ([1024, 6, 8], 1),
                                          ([512, 12, 8], 2), ([512, 130, 8], 1)])
def test_tma3d_matches_oracle(cm, oracle, shape, chains):
    n = nsites(shape)
    occs = [rand_occ(n, 23 + ch) for ch in range(chains)]
    conds = [(5235.0, 0.05), (4000.0, -0.03)][:chains]
    lat = run_cb(cm, shape, occs if chains > 1 else occs[0], None if chains > 1 else conds[0][0],
                 None if chains > 1 else conds[0][1], 99, 3,
                 "tma3d:js=44" if shape[1] == 130 else "tma3d",  # 130 columns: two strips of 64 and 66
                 n_chains=chains,
                 chain_conditions=conds if chains > 1 else None)
    assert lat.kernel_variant == "tma3d"
    for ch, (T, mu) in enumerate(conds):
        ref = oracle.checkerboard_run(shape, occs[ch], J, T, mu, 99, ch, 0, 3, 1)
        assert np.array_equal(lat.download(ch), ref["occupation"])
        S, B = lat.samples_sb(ch)
        assert np.array_equal(S, ref["S"]) and np.array_equal(B, ref["B"])
        assert lat.counters(ch)[1] == ref["n_accept"]


def test_tma3d_rejects_lattices_it_cannot_hold(cm):
    lat = cm.IsingLatticeGPU([256, 8, 8], J=J)
    lat.set_conditions(5235.0, 0.0)
    lat.set_kernel_variant("tma3d")
    with pytest.raises(cm.CmgError, match="tma3d needs"):
        lat.run_passes(1, cm.MODE_CHECKERBOARD, 0)


def test_seven_philox_rounds_opt_in_matches_the_oracle(cm, oracle):
    """cmg_set_philox_rounds(7): the seven-round stream (the fewest rounds of Philox4x32 that pass
    BigCrush) in the resident, the streaming 2-d and the generic kernel == the oracle's restatement
    with seven rounds; a different trajectory from the default; kernels without it refuse."""
    T, mu, seed = 2633.0, 0.013, 31337
    for shape, variants, n_passes in (([1024, 512], ("ring2d", "bulk2d", "generic", "auto"), 6), ([64, 48], ("bulk2d", "generic", "auto"), 5),
                                      ([32, 4, 6], ("generic", "auto"), 4)):
        occ = rand_occ(nsites(shape), 77)
        ref7 = oracle.checkerboard_run(shape, occ, J, T, mu, seed, 0, 0, n_passes, 2, philox_rounds=7)
        ref10 = oracle.checkerboard_run(shape, occ, J, T, mu, seed, 0, 0, n_passes, 2)
        assert not np.array_equal(ref7["occupation"], ref10["occupation"])
        for variant in variants:
            lat = cm.IsingLatticeGPU(shape, J=J)
            lat.set_conditions(T, mu)
            lat.seed_philox(seed)
            lat.set_philox_rounds(7)
            lat.set_kernel_variant(variant)
            lat.upload(occ)
            lat.run_passes(n_passes, cm.MODE_CHECKERBOARD, 2)
            assert np.array_equal(lat.download(), ref7["occupation"]), (shape, variant, lat.kernel_variant)
            S, B = lat.samples_sb()
            assert np.array_equal(S, ref7["S"]) and np.array_equal(B, ref7["B"])
            assert lat.counters()[1] == ref7["n_accept"]
            lat.close()
    lat = cm.IsingLatticeGPU([64, 48], J=J)
    lat.set_conditions(T, mu)
    with pytest.raises(cm.CmgError):
        lat.set_philox_rounds(8)
    lat.set_philox_rounds(7)
    lat.set_kernel_variant("tile2d")
    with pytest.raises(cm.CmgError, match="seven Philox rounds"):
        lat.run_passes(1, cm.MODE_CHECKERBOARD, 0)
    lat.close()


def test_multichain_grid_matches_oracle(cm, oracle):
    # a 2x3 (T, mu) grid of independent chains, one context (BASELINE config 4 in small)
    shape = [64, 32]
    n = nsites(shape)
    conds = [(t, m) for t in (1500.0, 4000.0) for m in (-0.2, 0.0, 0.2)]
    occs = [rand_occ(n, 100 + i) for i in range(len(conds))]
    for variant in ("generic", "bulk2d", "tile2d"):
        lat = run_cb(cm, shape, occs, None, None, 777, 5, variant, n_chains=len(conds), chain_conditions=conds)
        for ch, (t, m) in enumerate(conds):
            ref = oracle.checkerboard_run(shape, occs[ch], J, t, m, 777, ch, 0, 5, 1)
            assert np.array_equal(lat.download(ch), ref["occupation"])
            S, B = lat.samples_sb(ch)
            assert np.array_equal(S, ref["S"]) and np.array_equal(B, ref["B"])
            assert lat.counters(ch)[1] == ref["n_accept"]
            assert np.array_equal(lat.samples(cm.Q_POTENTIAL_ENERGY, ch), ref["potential_energy"])


def test_pass_counter_continuation(cm, oracle):
    # two run_passes calls continue the same Philox stream as one call
    shape = [64, 16]
    occ = rand_occ(nsites(shape), 31)
    a = run_cb(cm, shape, occ, 2500.0, 0.0, 5, 6, "bulk2d")
    b = run_cb(cm, shape, occ, 2500.0, 0.0, 5, 2, "tile2d")
    b.run_passes(4, cm.MODE_CHECKERBOARD, 1)
    assert np.array_equal(a.download(), b.download())
    assert np.array_equal(a.samples_sb()[1], b.samples_sb()[1])
    ref = oracle.checkerboard_run(shape, occ, J, 2500.0, 0.0, 5, 0, 3, 2, 1)  # pass0 = 3
    c = cm.IsingLatticeGPU(shape, J=J)
    c.set_conditions(2500.0, 0.0)
    c.seed_philox(5)
    c.set_pass_counter(3)
    c.upload(occ)
    c.run_passes(2, cm.MODE_CHECKERBOARD, 1)
    assert np.array_equal(c.download(), ref["occupation"])


# -------------------------------------------------------- serial reference ----
def test_device_rng_matches_libstdcxx(cm, oracle):
    lat = cm.IsingLatticeGPU([4, 4], J=J)
    lat.seed_mt19937_64(5489)
    e = oracle.RandomNumberEngine()
    e.seed(5489)
    reqs = []
    rng = np.random.default_rng(0)
    int_choices = [9, 624, 9999, 2**31, 2**64 - 1, 16777215]
    real_choices = [1.0, 9.0, 0.1]
    for i in range(700):  # crosses two state regenerations
        if i % 3 == 2:
            reqs.append(("real", real_choices[int(rng.integers(len(real_choices)))]))
        else:
            reqs.append(("int", int_choices[int(rng.integers(len(int_choices)))]))
    got = lat.rng_draw(reqs)
    for (kind, mx), g in zip(reqs, got):
        want = oracle.random_real(e, mx) if kind == "real" else oracle.random_int(e, mx)
        assert g == want
    # engine state round trip == operator<< of the host engine
    assert lat.engine_dump() == e.dump()
    e2 = oracle.RandomNumberEngine()
    e2.seed(42)
    lat.load_engine_dump(e2.dump())
    assert lat.rng_draw([("int", 9)] * 5) == [oracle.random_int(e2, 9) for _ in range(5)]


@pytest.mark.parametrize(
    "shape,T,mu,n_passes,seed",
    [([25, 25], 2000.0, 0.0, 40, 1), ([100, 100], 2633.0, 0.0, 10, 12345), ([6, 4], 800.0, 0.3, 50, 7), ([5, 4, 3], 4000.0, 0.05, 30, 3), ([512, 512], 2633.0, 0.0, 1, 9)],
)
def test_serial_mode_reproduces_reference_trajectory(cm, oracle, shape, T, mu, n_passes, seed):
    n = nsites(shape)
    occ = np.ones(n, dtype=np.int32) if shape[0] == 25 else rand_occ(n, 77)
    lat = cm.IsingLatticeGPU(shape, J=J)
    lat.set_conditions(T, mu)
    lat.seed_mt19937_64(seed)
    lat.upload(occ)
    lat.run_passes(n_passes, cm.MODE_SERIAL_REFERENCE, sample_period=1)
    e = oracle.RandomNumberEngine()
    e.seed(seed)
    ref = oracle.sgc_run(shape, occ, J, T, mu, True, e, {"max_count": n_passes}, 1)
    assert np.array_equal(lat.download(), ref["occupation"])
    n_pass, n_acc, n_rej = lat.counters()
    assert (n_pass, n_acc, n_rej) == (ref["n_pass"], ref["n_accept"], ref["n_reject"])
    assert np.array_equal(lat.samples(cm.Q_PARAM_COMPOSITION), ref["samplers"]["param_composition"])
    assert np.array_equal(lat.samples(cm.Q_FORMATION_ENERGY), ref["samplers"]["formation_energy"])
    assert np.array_equal(lat.samples(cm.Q_POTENTIAL_ENERGY), ref["samplers"]["potential_energy"])
    assert lat.engine_dump() == e.dump()  # same number of draws consumed


def test_serial_mode_multichain_and_continuation(cm, oracle):
    shape = [10, 8]
    occ = rand_occ(80, 3)
    lat = cm.IsingLatticeGPU(shape, n_chains=3, J=J)
    conds = [(1000.0, 0.0), (2633.0, 0.1), (6000.0, -0.1)]
    for ch, (t, m) in enumerate(conds):
        lat.set_conditions(t, m, chain=ch)
        lat.seed_mt19937_64(100 + ch, chain=ch)
        lat.upload(occ, ch)
    lat.run_passes(7, cm.MODE_SERIAL_REFERENCE, sample_period=2)
    lat.run_passes(5, cm.MODE_SERIAL_REFERENCE, sample_period=2)
    for ch, (t, m) in enumerate(conds):
        e = oracle.RandomNumberEngine()
        e.seed(100 + ch)
        ref = oracle.sgc_run(shape, occ, J, t, m, True, e, {"max_count": 12}, 2)
        assert np.array_equal(lat.download(ch), ref["occupation"])
        assert np.array_equal(lat.samples(cm.Q_POTENTIAL_ENERGY, ch), ref["samplers"]["potential_energy"])
        assert lat.counters(ch)[1] == ref["n_accept"]


@pytest.mark.parametrize("shape,variant", [([64, 48], "auto"), ([96, 34], "bulk2d"), ([1024, 96], "ring2d"), ([32, 6], "generic")])
def test_nonlist_energy_form_sampled_on_the_device(cm, oracle, shape, variant):
    # use_nlist = false (model.hh:273-285): rows then columns, -J * dot each, every
    # term rounded -- a J that is not dyadic makes the two forms differ in the last bits
    n = nsites(shape)
    occ = rand_occ(n, 31)
    Jx, T, mu = 0.1, 2633.0, 0.03
    lat = cm.IsingLatticeGPU(shape, J=Jx)
    lat.set_conditions(T, mu)
    lat.seed_philox(99)
    lat.set_kernel_variant(variant)
    lat.set_energy_form(use_nlist=False)
    lat.upload(occ)
    n_passes, period = 9, 2
    lat.run_passes(n_passes, cm.MODE_CHECKERBOARD, period)
    ef = lat.samples(cm.Q_FORMATION_ENERGY)
    ep = lat.samples(cm.Q_POTENTIAL_ENERGY)
    x = lat.samples(cm.Q_PARAM_COMPOSITION)
    assert len(ef) == n_passes // period
    cur, differs = occ, False
    for k in range(n_passes // period):
        ref = oracle.checkerboard_run(shape, cur, Jx, T, mu, 99, 0, k * period, period, period)
        cur = ref["occupation"]
        assert ef[k] == oracle.formation_energy(shape, cur, Jx, False)[1]
        assert ep[k] == oracle.potential(shape, cur, Jx, T, mu, False)[1]
        assert x[k] == ref["param_composition"][0]
        differs |= ef[k] != ref["formation_energy"][0]
    rest = n_passes % period  # passes after the last sample
    if rest:
        cur = oracle.checkerboard_run(shape, cur, Jx, T, mu, 99, 0, n_passes - rest, rest, 0)["occupation"]
    assert np.array_equal(lat.download(), cur)
    # statistics read the same columns
    st = lat.series_stats(cm.Q_FORMATION_ENERGY, 0)
    assert math.isclose(st["mean"], float(np.mean(ef)), rel_tol=1e-12)
    # switching the form needs an empty series
    with pytest.raises(cm.CmgError, match="empty sample series"):
        lat.set_energy_form(use_nlist=True)
    lat.clear_samples()
    lat.set_energy_form(use_nlist=True)
    lat.run_passes(1, cm.MODE_CHECKERBOARD, 2)  # pass 10 of the schedule: sampled
    assert lat.samples(cm.Q_FORMATION_ENERGY)[0] == oracle.formation_energy(shape, lat.download(), Jx, True)[1]


def test_nonlist_energy_form_is_refused_where_unsupported(cm):
    lat = cm.IsingLatticeGPU([16, 16, 16], J=J)
    with pytest.raises(cm.CmgError, match="use_nlist=false"):
        lat.set_energy_form(use_nlist=False)


def test_device_reproduces_the_golden_vectors(cm):
    # tests/golden/ising_sgc_golden.json (made by tests/golden/make_golden.py): the
    # device against frozen vectors directly, without the oracle in the loop
    import hashlib
    import json
    import os

    from casmcode_monte_b200.lattice import host_series_equilibration, host_series_equilibration_weighted, host_series_stats, host_series_stats_weighted

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ising_sgc_golden.json")) as f:
        g = json.load(f)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    unhex = lambda v: np.array([float.fromhex(s) for s in v])
    for c in g["checkerboard"]:
        n = nsites(c["shape"])
        occ = np.random.default_rng(c["occ_seed"]).choice(np.array([-1, 1], dtype=np.int32), size=n)
        variants = ["auto", "generic"] + (["ring2d"] if c["shape"] == [1024, 128] else [])
        for variant in variants:
            lat = run_cb(cm, c["shape"], occ, c["T"], c["mu"], c["philox_seed"], c["n_passes"], variant, sample_period=c["sample_period"])
            assert sha(lat.download().astype(np.int32)) == c["occupation_sha256"], (c["shape"], variant)
            S, B = lat.samples_sb()
            assert [int(v) for v in S] == c["S"] and [int(v) for v in B] == c["B"]
            assert lat.counters()[1] == c["n_accept"]
            assert np.array_equal(lat.samples(cm.Q_POTENTIAL_ENERGY), unhex(c["potential_energy"]))
            assert np.array_equal(lat.samples(cm.Q_PARAM_COMPOSITION), unhex(c["param_composition"]))
    for c in g["serial"]:
        lat = cm.IsingLatticeGPU(c["shape"], J=g["J"])
        lat.set_conditions(c["T"], c["mu"])
        lat.seed_mt19937_64(c["mt19937_64_seed"])
        lat.upload(np.ones(nsites(c["shape"]), dtype=np.int32))
        lat.run_passes(c["max_count"], cm.MODE_SERIAL_REFERENCE, 1)
        assert sha(lat.download().astype(np.int32)) == c["occupation_sha256"]
        assert lat.counters()[1] == c["n_accept"]
        assert np.array_equal(lat.samples(cm.Q_POTENTIAL_ENERGY), unhex(c["potential_energy"]))
    for c in g["statistics"]:
        x, w = unhex(c["x"]), unhex(c["w"])
        st = host_series_stats(x)
        assert st["k_star"] == c["k_star"]
        assert math.isclose(st["mean"], float.fromhex(c["mean"]), rel_tol=1e-12)
        assert math.isclose(st["calculated_precision"], float.fromhex(c["precision"]), rel_tol=1e-10)
        assert list(host_series_equilibration(x, 0.05)) == c["equilibration_abs_0.05"]
        assert list(host_series_equilibration_weighted(x, w, 0.05)) == c["weighted_equilibration_abs_0.05"]
        for method, key in ((1, "weighted_method1"), (2, "weighted_method2")):
            sw = host_series_stats_weighted(x, w, method=method, n_resamples=1000)
            assert math.isclose(sw["mean"], float.fromhex(c[key][0]), rel_tol=1e-12)
            assert math.isclose(sw["calculated_precision"], float.fromhex(c[key][1]), rel_tol=1e-10)


# -------------------------------------------------------------- statistics ----
def test_series_statistics_and_equilibration(cm, oracle):
    shape = [64, 64]
    lat = cm.IsingLatticeGPU(shape, n_chains=2, J=J)
    lat.set_conditions(2400.0, 0.0, chain=0)
    lat.set_conditions(3000.0, 0.05, chain=1)
    lat.seed_philox(3)
    lat.run_passes(1500, cm.MODE_CHECKERBOARD, 1)
    for ch in range(2):
        for q in (cm.Q_PARAM_COMPOSITION, cm.Q_POTENTIAL_ENERGY):
            x = lat.samples(q, ch)
            eq_ref = oracle.default_equilibration_check(x, abs=1e-3)
            assert lat.series_equilibration(q, 1e-3, ch) == eq_ref
            first = eq_ref[1] if eq_ref[0] else 0
            st = lat.series_stats(q, ch, first=first)
            mean, prec = oracle.basic_statistics(x[first:])
            f, k = oracle.autocorrelation_factor(x[first:])
            assert st["k_star"] == k
            assert math.isclose(st["mean"], mean, rel_tol=1e-12)
            assert math.isclose(st["calculated_precision"], prec, rel_tol=1e-10)
    m, p, v, k = lat.series_stats_all(cm.Q_POTENTIAL_ENERGY)
    e, n = lat.series_equilibration_all(cm.Q_POTENTIAL_ENERGY, 1e-3)
    for ch in range(2):
        x = lat.samples(cm.Q_POTENTIAL_ENERGY, ch)
        assert math.isclose(m[ch], oracle.basic_statistics(x)[0], rel_tol=1e-12)
        assert (bool(e[ch]), int(n[ch])) == oracle.default_equilibration_check(x, abs=1e-3)


def test_host_series_statistics_edge_cases(cm, oracle):
    from casmcode_monte_b200.lattice import host_series_equilibration, host_series_stats

    rng = np.random.default_rng(1)
    cases = [
        np.full(50, 2.5),  # no variation -> f = 1 (BasicStatistics.cc:31-33)
        np.arange(10, dtype=float),
        rng.normal(size=1),
        rng.normal(size=2),
        rng.normal(size=1001) + 5,
        np.cumsum(rng.normal(size=3000)) * 0.05 + rng.normal(size=3000),
        np.concatenate([np.linspace(5, 0, 50), np.zeros(450)]) + rng.normal(scale=0.01, size=500),
    ]
    for x in cases:
        st = host_series_stats(x)
        mean, prec = oracle.basic_statistics(x)
        f, k = oracle.autocorrelation_factor(x)
        assert st["k_star"] == k
        assert math.isclose(st["mean"], mean, rel_tol=1e-12, abs_tol=1e-300)
        if math.isfinite(prec) and prec < 1e300:
            assert math.isclose(st["calculated_precision"], prec, rel_tol=1e-10, abs_tol=1e-300)
        else:
            assert not (st["calculated_precision"] < 1e300)
        for p in (1e-3, 0.05):
            assert host_series_equilibration(x, p) == oracle.default_equilibration_check(x, abs=p)


def test_weighted_statistics_match_oracle(cm, oracle):
    # BasicStatistics.cc:50-73, :144-188; EquilibrationCheck.cc:137-161.  The
    # resampling walk selects samples by floating-point comparisons, so it must
    # be identical; the reductions are tree-ordered on the device (tolerances).
    from casmcode_monte_b200.lattice import (
        host_series_equilibration_weighted,
        host_series_resample,
        host_series_stats_weighted,
    )

    rng = np.random.default_rng(11)
    cases = []
    for n in (1, 2, 37, 1000, 5000):
        x = np.cumsum(rng.normal(size=n)) * 0.1 + rng.normal(size=n) + 3.0
        w = rng.exponential(size=n) + 1e-3  # residence times of a rejection-free walk
        cases.append((x, w))
    cases.append((np.full(200, 1.5), rng.exponential(size=200)))  # no variation
    cases.append((rng.normal(size=300), np.ones(300)))  # equal weights
    x = np.concatenate([np.linspace(4, 0, 100), np.zeros(900)]) + rng.normal(scale=0.02, size=1000)
    cases.append((x, rng.uniform(0.5, 1.5, size=1000)))
    for x, w in cases:
        W = 0.0  # summed in order, as the oracle and the device do
        for v in w:
            W += float(v)
        for R in (10, 1000, 10000):
            assert np.array_equal(host_series_resample(x, w, W, R), oracle.resample(x, w, W, R))
            for method in (1, 2):
                st = host_series_stats_weighted(x, w, method=method, n_resamples=R)
                mean, prec = oracle.basic_statistics(x, w, method=method, n_resamples=R)
                assert st["weight_sum"] == W
                assert math.isclose(st["mean"], mean, rel_tol=1e-12, abs_tol=1e-300)
                if math.isfinite(prec) and prec < 1e300:
                    # abs_tol: a constant series has a variance of rounding noise only
                    assert math.isclose(st["calculated_precision"], prec, rel_tol=1e-10, abs_tol=1e-13 * max(1.0, abs(mean)))
                else:
                    assert not (st["calculated_precision"] < 1e300)
        for p in (1e-3, 0.05, 0.5):
            assert host_series_equilibration_weighted(x, w, p) == oracle.default_equilibration_check(x, w, abs=p)


def test_weighted_statistics_through_the_mirrored_classes(cm, oracle):
    from casmcode_monte_b200.monte import sampling

    rng = np.random.default_rng(5)
    x = np.cumsum(rng.normal(size=800)) * 0.05 + rng.normal(size=800)
    w = rng.exponential(size=800) + 1e-3
    for method in (1, 2):
        calc = sampling.BasicStatisticsCalculator(confidence=0.9, weighted_observations_method=method, n_resamples=2000)
        st = calc(x, w)
        mean, prec = oracle.basic_statistics(x, w, confidence=0.9, method=method, n_resamples=2000)
        assert math.isclose(st.mean, mean, rel_tol=1e-12)
        assert math.isclose(st.calculated_precision, prec, rel_tol=1e-10)
    calc = sampling.BasicStatisticsCalculator(weighted_observations_method=3)
    with pytest.raises(RuntimeError, match="invalid method"):
        calc(x, w)
    with pytest.raises(RuntimeError, match="observations.size\\(\\) != sample_weight.size"):
        sampling.BasicStatisticsCalculator()(x, w[:-1])
    rp = sampling.RequestedPrecision(abs=0.05)
    r = sampling.default_equilibration_check(x, w, rp)
    assert (r.is_equilibrated, r.N_samples_for_equilibration) == oracle.default_equilibration_check(x, w, abs=0.05)

    # a weighted completion check: Sampler columns + a sample_weight Sampler
    s = sampling.Sampler(shape=[])
    sw = sampling.Sampler(shape=[])
    for xi, wi in zip(x, w):
        s.append(np.array([xi]))
        sw.append(np.array([wi]))
    assert np.array_equal(sw.component(0), w)


# ------------------------------------------------------------- conversions ----
def test_conversions_batch(cm, oracle):
    from casmcode_monte_b200.lattice import conv_bijk_to_l, conv_l_to_bijk

    n3, nb = [3, 4, 5], 2
    l = np.arange(3 * 4 * 5 * 2)
    bijk = conv_l_to_bijk(n3, nb, l)
    for li in (0, 1, 59, 60, 119):
        assert tuple(bijk[li]) == oracle.conv_l_to_bijk(n3, nb, int(li))
    assert np.array_equal(conv_bijk_to_l(n3, nb, bijk), l)
    # python/tests/events/test_Conversions.py:65-72 : l = b*N_unitcells + index, periodic wrap
    assert conv_bijk_to_l([3, 3, 3], 2, [[1, 0, 0, 0]])[0] == 27
    assert tuple(conv_l_to_bijk([3, 3, 3], 2, [27])[0]) == (1, 0, 0, 0)
    wrapped = conv_bijk_to_l([3, 3, 3], 2, [[0, 3, -1, 4], [1, -3, 5, -7]])
    assert list(wrapped) == [oracle.conv_bijk_to_l([3, 3, 3], 2, 0, 3, -1, 4), oracle.conv_bijk_to_l([3, 3, 3], 2, 1, -3, 5, -7)]
    with pytest.raises(cm.CmgError):
        conv_l_to_bijk([3, 3, 3], 2, [54])


def test_general_conversions_batched_on_the_device_match_the_host():
    import casmcode_monte_b200.monte.events as ev

    for T in ([[-1, 1, 1], [1, -1, 1], [1, 1, -1]], [[2, 0, 0], [0, 3, 0], [0, 0, 1]], [[2, 1, 0], [0, 3, 1], [1, 0, 2]], [[5, 0, 0], [0, 5, 0], [0, 0, 5]]):
        c = ev.Conversions(occ_dof=[["A", "B"]] * 3, transformation_matrix_to_super=np.array(T))
        ls = list(range(c.l_size()))
        bijk = c.l_to_bijk_batch(ls)
        assert [list(map(int, r)) for r in bijk] == [c.l_to_bijk(l) for l in ls]
        shifted = bijk.copy()
        shifted[:, 1:] += (np.array(T) @ np.array([2, -1, 3]))[None, :]  # a supercell translation
        assert list(c.bijk_to_l_batch(shifted)) == ls


# (512 x 96 / 768 x 64 through tile2d: tiles with halos write back to the second copy of the planes and
# the two are swapped after every launch -- 7 passes are three launches, an odd number of swaps)
@pytest.mark.parametrize("shape,variant", [([64, 48], "auto"), ([1024, 128], "ring2d"), ([256, 64], "bulk2d"), ([32, 10, 8], "auto"),
                                           ([512, 96], "tile2d"), ([768, 64], "tile2d:p=2"), ([4096, 64], "bulk2d")])
def test_mark_and_rollback_undo_a_speculative_block(cm, oracle, shape, variant):
    """cmg_mark / cmg_rollback: passes enqueued after the mark leave no trace after a rollback
    (occupation, acceptance count, pass and sample counters, sample series), the samples up to
    the mark can be checked meanwhile, and the trajectory continues as if nothing happened."""
    n = nsites(shape)
    occ = rand_occ(n, 77)
    T, mu, seed = 2633.0 if len(shape) == 2 else 5235.0, 0.02, 4711
    lat = cm.IsingLatticeGPU(shape, J=J)
    lat.set_conditions(T, mu)
    lat.seed_philox(seed)
    lat.set_kernel_variant(variant)
    lat.upload(occ)
    lat.run_passes(5, cm.MODE_CHECKERBOARD, 1)
    lat.mark()
    lat.run_passes(7, cm.MODE_CHECKERBOARD, 1)  # speculative
    chk = lat.series_check((cm.Q_POTENTIAL_ENERGY, cm.Q_PARAM_COMPOSITION), [1e-3, 1e-3], count=5)  # next to it
    lat.rollback()
    ref5 = oracle.checkerboard_run(shape, occ, J, T, mu, seed, 0, 0, 5, 1)
    assert np.array_equal(lat.download(), ref5["occupation"])
    assert lat.counters()[:2] == (5, ref5["n_accept"]) and lat.n_samples == 5
    S, B = lat.samples_sb()
    assert np.array_equal(S, ref5["S"]) and np.array_equal(B, ref5["B"])
    assert np.array_equal(lat.samples(cm.Q_POTENTIAL_ENERGY), ref5["potential_energy"])
    same = lat.series_check((cm.Q_POTENTIAL_ENERGY, cm.Q_PARAM_COMPOSITION), [1e-3, 1e-3], count=5)
    assert chk == same
    lat.run_passes(7, cm.MODE_CHECKERBOARD, 1)
    ref12 = oracle.checkerboard_run(shape, occ, J, T, mu, seed, 0, 0, 12, 1)
    assert np.array_equal(lat.download(), ref12["occupation"])
    S, B = lat.samples_sb()
    assert np.array_equal(S, ref12["S"]) and np.array_equal(B, ref12["B"]) and lat.counters()[1] == ref12["n_accept"]
    with pytest.raises(cm.CmgError):
        lat.rollback()  # no mark any more
    lat.close()


# ------------------------------------- full-size, size-independent properties ----
def test_full_size_4096_properties(cm):
    shape = [4096, 4096]
    n = nsites(shape)
    lat = cm.IsingLatticeGPU(shape, n_chains=2, J=J)
    lat.set_conditions(2633.0, 0.0)
    lat.seed_philox(0xC0FFEE)
    lat.randomize(12345, 0.5, chain=0)
    a0 = lat.download(0)
    assert abs(a0.mean()) < 5e-3 and set(np.unique(a0)) == {-1, 1}
    lat.upload(-a0, 1)  # chain 1 = spin-inverted copy; mu = 0 => Z2-symmetric dynamics...
    # ...but chains use different Philox counters, so compare chain 0 of two contexts instead
    lat2 = cm.IsingLatticeGPU(shape, J=J)
    lat2.set_conditions(2633.0, 0.0)
    lat2.seed_philox(0xC0FFEE)
    lat2.upload(-a0)
    lat.run_passes(3, cm.MODE_CHECKERBOARD, 1)
    lat2.run_passes(3, cm.MODE_CHECKERBOARD, 1)
    a3 = lat.download(0)
    assert np.array_equal(lat2.download(0), -a3)  # Z2 symmetry at mu = 0
    S, B = lat.samples_sb(0)
    S2, B2 = lat2.samples_sb(0)
    assert np.array_equal(S, -S2) and np.array_equal(B, B2)
    # fused sampling == independent reduction of the final state == numpy on the download
    assert lat.sample_now(0) == (int(S[-1]), int(B[-1]))
    g = a3.reshape(4096, 4096, order="F").astype(np.int64)
    assert int(g.sum()) == S[-1]
    assert int((g * (np.roll(g, -1, 0) + np.roll(g, -1, 1))).sum()) == B[-1]
    assert lat.counters(0)[1] + lat.counters(0)[2] == 3 * n
    # the tiled kernel (auto), bulk2d and the generic kernel agree at full size
    assert lat.kernel_variant in ("tile2d", "ring2d")
    for variant in ("generic", "bulk2d", "tile2d", "ring2d"):
        lat3 = cm.IsingLatticeGPU(shape, J=J)
        lat3.set_conditions(2633.0, 0.0)
        lat3.seed_philox(0xC0FFEE)
        lat3.set_kernel_variant(variant)
        lat3.upload(a0)
        lat3.run_passes(3, cm.MODE_CHECKERBOARD, 1)
        assert np.array_equal(lat3.download(0), a3)
        assert np.array_equal(lat3.samples_sb(0)[1], B)
    # frozen limit: T -> 0+ from all-up never flips (dE = 8J > 0, exp(-dE*beta) underflows)
    lat4 = cm.IsingLatticeGPU(shape, J=J)
    lat4.set_conditions(1.0, 0.0)
    lat4.run_passes(2, cm.MODE_CHECKERBOARD, 1)
    assert lat4.counters(0)[1] == 0 and lat4.sample_now(0) == (n, 2 * n)


def test_3d_full_plane_properties(cm):
    shape = [128, 64, 32]
    n = nsites(shape)
    lat = cm.IsingLatticeGPU(shape, J=J)
    lat.set_conditions(5235.0, 0.0)
    lat.seed_philox(11)
    lat.randomize(5, 0.5)
    a0 = lat.download()
    lat.run_passes(2, cm.MODE_CHECKERBOARD, 1)
    a2 = lat.download()
    lat2 = cm.IsingLatticeGPU(shape, J=J)
    lat2.set_conditions(5235.0, 0.0)
    lat2.seed_philox(11)
    lat2.set_kernel_variant("generic")
    lat2.upload(a0)
    lat2.run_passes(2, cm.MODE_CHECKERBOARD, 1)
    assert np.array_equal(lat2.download(), a2)
    g = a2.reshape(shape, order="F").astype(np.int64)
    S, B = lat.samples_sb()
    assert int(g.sum()) == S[-1]
    assert int((g * (np.roll(g, -1, 0) + np.roll(g, -1, 1) + np.roll(g, -1, 2))).sum()) == B[-1]
    assert lat.sample_now() == (int(S[-1]), int(B[-1]))


# ------------------------------------------------- chained half-sweeps ----
# Consecutive half-sweeps of a run are launched as programmatic dependents and wait for the
# neighbour CTAs of the previous half-sweep only (per-CTA flags) -- half-sweeps overlap in
# time.  The trajectory, the sampled sums and the counters must equal those of the plain
# one-grid-after-the-other launches bit for bit, on lattices large enough for the CTAs of a
# half-sweep to drift apart (several waves, several chains, 2-d and 3-d, V below / at / above
# one CTA per strip).
@pytest.mark.parametrize(
    "shape,chains,variant,n_passes",
    [([4096, 4096], 2, "bulk2d", 24), ([1024, 2048], 3, "bulk2d", 40), ([8192, 1024], 1, "bulk2d", 30),
     ([256, 4096], 2, "bulk2d:js=16", 30), ([512, 128, 128], 1, "bulk3d", 20), ([256, 64, 96], 2, "bulk3d", 30),
     ([1024, 64, 32], 1, "bulk3d", 20), ([4096, 16, 8], 1, "bulk3d", 20), ([64, 40, 24], 1, "bulk3d", 30)])
def test_chained_half_sweeps_equal_whole_grid_launches(cm, shape, chains, variant, n_passes):
    states = []
    for opts in ("", ":pdl=0"):
        lat = cm.IsingLatticeGPU(shape, n_chains=chains, J=J)
        for ch in range(chains):
            lat.set_conditions(2633.0 if len(shape) == 2 else 5235.0, 0.01 * ch, chain=ch)
            lat.randomize(77 + ch, 0.5, chain=ch)
        lat.seed_philox(0xABCDEF)
        lat.set_kernel_variant(variant + opts)
        lat.run_passes(n_passes, cm.MODE_CHECKERBOARD, 3)
        lat.run_passes(5, cm.MODE_CHECKERBOARD, 1)  # a second call: its first half-sweep waits for the whole grid
        lat.sync()
        states.append(([lat.download(ch) for ch in range(chains)], [lat.samples_sb(ch) for ch in range(chains)],
                       [lat.counters(ch) for ch in range(chains)]))
        lat.close()
    (occ_a, sb_a, cnt_a), (occ_b, sb_b, cnt_b) = states
    for ch in range(chains):
        assert np.array_equal(occ_a[ch], occ_b[ch])
        assert np.array_equal(sb_a[ch][0], sb_b[ch][0]) and np.array_equal(sb_a[ch][1], sb_b[ch][1])
        assert cnt_a[ch] == cnt_b[ch]


# One lattice per CTA with warps that wait for their two neighbour warps instead of for the
# CTA (k_tile2d, single tile): many passes per launch, several chains, column groups of 4 /
# 8 / 16 / 32 threads and CTAs of 256 / 512 / 1024 threads, against the oracle.
@pytest.mark.parametrize(
    "shape,variant",
    [([256, 256], "tile2d"), ([256, 256], "tile2d:nt=1024"), ([256, 256], "tile2d:nt=256"), ([128, 256], "tile2d"),
     ([512, 128], "tile2d"), ([1024, 64], "tile2d"), ([256, 70], "tile2d")])
def test_tile2d_neighbour_warp_waits_match_oracle(cm, oracle, shape, variant):
    n = nsites(shape)
    conds = [(2633.0, 0.0), (2200.0, 0.03), (3500.0, -0.05)]
    occs = [rand_occ(n, 300 + i) for i in range(len(conds))]
    lat = run_cb(cm, shape, occs, None, None, 99, 60, variant, n_chains=len(conds), chain_conditions=conds, sample_period=7)
    assert lat.kernel_variant == "tile2d"
    for ch, (t, m) in enumerate(conds):
        ref = oracle.checkerboard_run(shape, occs[ch], J, t, m, 99, ch, 0, 60, 7)
        assert np.array_equal(lat.download(ch), ref["occupation"])
        S, B = lat.samples_sb(ch)
        assert np.array_equal(S, ref["S"]) and np.array_equal(B, ref["B"])
        assert lat.counters(ch)[1] == ref["n_accept"]


# The resident kernel over many half-sweeps per launch: warps wait for their neighbour warps
# on per-warp mbarriers (two per warp, alternating with the colour), one / two / four / eight
# warps per column group, against the oracle.
# (n0 = 256 / 512: the 128-thread form, column groups of 8 / 16 threads, several groups per warp)
@pytest.mark.parametrize("shape,variant", [([1024, 256], "ring2d"), ([2048, 160], "ring2d"), ([4096, 300], "ring2d:rp=17"),
                                           ([8192, 64], "ring2d"), ([512, 512], "ring2d"), ([256, 1024], "ring2d"), ([512, 70], "ring2d:rp=9"),
                                           ([256, 100], "ring2d"), ([512, 2048], "ring2d")])
def test_ring2d_neighbour_warp_waits_over_many_passes(cm, oracle, shape, variant):
    n = nsites(shape)
    occ = rand_occ(n, 77)
    T, mu = 2633.0, 0.01
    lat = run_cb(cm, shape, occ, T, mu, 31337, 40, variant, sample_period=3)
    assert lat.kernel_variant == "ring2d"
    ref = oracle.checkerboard_run(shape, occ, J, T, mu, 31337, 0, 0, 40, 3)
    assert np.array_equal(lat.download(), ref["occupation"])
    S, B = lat.samples_sb()
    assert np.array_equal(S, ref["S"]) and np.array_equal(B, ref["B"])
    assert lat.counters()[1] == ref["n_accept"]


def test_statistics_fuzz_against_the_oracle():
    """tools/stats_fuzz.py: 250 random series (constant, drifting, stepping, heavy-tailed, discrete,
    one to 6000 samples, lengths around the chunk sizes of the equilibration scan) through the device
    statistics and the oracle."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "stats_fuzz.py"), "250", "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_host_formats_after_tiled_launches(cm, oracle):
    """int32 / int8 / bit-packed downloads, an upload in between and the natural-layout observables
    see the planes the last tiled launch wrote (tiles with halos alternate between two copies)."""
    shape = [512, 96]
    n = nsites(shape)
    occ = rand_occ(n, 5)
    T, mu, seed = 2633.0, 0.01, 99
    lat = cm.IsingLatticeGPU(shape, J=J)
    lat.set_conditions(T, mu)
    lat.seed_philox(seed)
    lat.set_kernel_variant("tile2d")
    lat.upload(occ)
    ref = occ
    done = 0
    for n_passes in (1, 3, 2, 5):  # 1, 1, 1 and 2 launches: the current copy alternates
        lat.run_passes(n_passes, cm.MODE_CHECKERBOARD, 1)
        r = oracle.checkerboard_run(shape, ref, J, T, mu, seed, 0, done, n_passes, 1)
        ref, done = r["occupation"], done + n_passes
        a32 = lat.download()
        assert np.array_equal(a32, ref)
        assert np.array_equal(lat.download_i8().astype(np.int32), ref)
        bits = lat.download_bits()
        assert np.array_equal(np.unpackbits(bits, bitorder="little")[:n].astype(np.int32) * 2 - 1, ref)
        assert lat.sample_now() == (int(r["S"][-1]), int(r["B"][-1]))
    other = rand_occ(n, 6)
    lat.upload_i8(other.astype(np.int8))
    assert np.array_equal(lat.download(), other)
    lat.run_passes(3, cm.MODE_CHECKERBOARD, 1)
    r = oracle.checkerboard_run(shape, other, J, T, mu, seed, 0, done, 3, 1)
    assert np.array_equal(lat.download(), r["occupation"])
    lat.close()
