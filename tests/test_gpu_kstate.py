"""k-state lattice model behind the general multi-species proposal machinery
(SURVEY 8f rank 3): OccCandidateList / OccLocation / propose_semigrand_canonical_event
(include/casm/monte/events/OccEventProposal.hh:260-348, src/casm/monte/events/
OccLocation.cc:39-116, :253-283, src/casm/monte/events/OccCandidate.cc:32-157).
The device paths against the line-by-line CPU restatement (oracle/kstate_oracle.hh):
tables and serial-order trajectories bit-identical, the checkerboard kernel
bit-identical to its scalar statement, and checkerboard vs serial ensemble
averages within 3 sigma."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

J = 0.1
V3 = np.array([[-J, J, 0.0], [J, -J, 0.5 * J], [0.0, 0.5 * J, -0.3 * J]])
MU3 = np.array([0.0, 0.02, -0.01])
V4 = np.array([[-J, 0.2 * J, 0.0, J], [0.2 * J, -0.5 * J, 0.3 * J, 0.0], [0.0, 0.3 * J, -J, 0.1 * J], [J, 0.0, 0.1 * J, -0.2 * J]])
MU4 = np.array([0.0, 0.01, -0.02, 0.03])


@pytest.fixture(scope="module")
def cm():
    import casmcode_monte_b200 as m

    return m


def make(cm, shape, V, T, mu, n_chains=1):
    lat = cm.IsingLatticeGPU(shape, n_chains=n_chains)
    lat.kstate_set_model(V)
    lat.kstate_set_conditions(T, mu)
    return lat


@pytest.mark.parametrize("dim,V,mu", [(2, V3, MU3), (3, V3, MU3), (2, V4, MU4), (2, np.array([[-J, J], [J, -J]]), np.array([0.0, 0.05]))])
def test_tables_bit_exact(cm, oracle, dim, V, mu):
    K = V.shape[0]
    lat = make(cm, [4, 4] if dim == 2 else [4, 4, 4], V, 2500.0, mu)
    dPhi, prob, thr, never = lat.kstate_tables()
    ref = oracle.kstate_table(dim, K, V, 2500.0, mu)
    assert dPhi.size == K * K * ref["n_cfg"]
    assert np.array_equal(dPhi, ref["dPhi"]) and np.array_equal(prob, ref["prob"])
    assert np.array_equal(thr, ref["thr_m1"]) and np.array_equal(never, ref["never"])


@pytest.mark.parametrize("shape,V,mu,T", [([16, 12], V3, MU3, 2500.0), ([9, 7], V3, MU3, 1200.0), ([8, 6, 4], V3, MU3, 4000.0), ([12, 10], V4, MU4, 3000.0)])
def test_serial_mode_reproduces_the_restated_proposal_machinery(cm, oracle, shape, V, mu, T):
    """Trajectory, counters, samples and engine state equal the restated loop: swap choice by
    cumulative candidate counts, choose_mol in the OccLocation list, Metropolis acceptance, apply."""
    K = V.shape[0]
    n = int(np.prod(shape))
    occ = np.random.default_rng(n).integers(0, K, size=n).astype(np.int32)
    lat = make(cm, shape, V, T, mu)
    lat.kstate_upload(occ)
    lat.seed_mt19937_64(4242)
    # two calls: the OccLocation lists persist between them
    lat.kstate_run_passes(7, cm.MODE_SERIAL_REFERENCE, 3)
    lat.kstate_run_passes(5, cm.MODE_SERIAL_REFERENCE, 3)
    e = oracle.RandomNumberEngine()
    e.seed(4242)
    ref = oracle.kstate_serial_run(shape, occ, K, V, T, mu, e, 12, 3)
    assert np.array_equal(lat.kstate_download(), ref["occupation"])
    n_pass, n_acc, n_rej = lat.counters()
    assert (n_pass, n_acc, n_rej) == (12, ref["n_accept"], ref["n_reject"])
    counts, bonds = lat.kstate_samples()
    assert np.array_equal(counts, ref["counts"]) and np.array_equal(bonds, ref["bonds"])
    st, pos = lat.get_engine_state()
    assert e.dump().split()[:312] == [str(int(v)) for v in st] and int(e.dump().split()[312]) == pos


# (odd numbers of plane indices per column: the last group of a column holds one site;
# K = 4 in 3-d: the threshold table is read from global memory; 300 x 70: several blocks along
# the column with a ragged last one, more columns than one block walks)
@pytest.mark.parametrize("shape,V,mu,T", [([16, 12], V3, MU3, 2500.0), ([64, 48], V4, MU4, 3000.0), ([8, 6, 4], V3, MU3, 4000.0),
                                          ([6, 4], V3, MU3, 2500.0), ([10, 6, 4], V3, MU3, 4000.0), ([8, 6, 4], V4, MU4, 4000.0),
                                          ([300, 70], V3, MU3, 2500.0), ([2, 2], V3, MU3, 2500.0)])
def test_checkerboard_kernel_matches_its_scalar_statement(cm, oracle, shape, V, mu, T):
    K = V.shape[0]
    n = int(np.prod(shape))
    occ = np.random.default_rng(7 + n).integers(0, K, size=n).astype(np.int32)
    lat = make(cm, shape, V, T, mu)
    lat.kstate_upload(occ)
    lat.seed_philox(0xC0FFEE)
    lat.kstate_run_passes(6, cm.MODE_CHECKERBOARD, 2)
    ref = oracle.kstate_checkerboard_run(shape, occ, K, V, T, mu, 0xC0FFEE, 0, 0, 6, 2)
    assert np.array_equal(lat.kstate_download(), ref["occupation"])
    assert lat.counters()[1] == ref["n_accept"]
    counts, bonds = lat.kstate_samples()
    assert np.array_equal(counts, ref["counts"]) and np.array_equal(bonds, ref["bonds"])
    assert counts.sum(axis=1).tolist() == [n] * 3 and bonds.sum(axis=(1, 2)).tolist() == [len(shape) * n] * 3


def test_checkerboard_ensemble_matches_serial_reference_within_3_sigma(cm, oracle):
    """3-state model, 24 independent chains per update order on 32 x 32: species fractions and the
    potential per site agree within 3 sigma (delete-one-chain jackknife over the chains)."""
    shape, K, T, M, n_eq, n_meas = [32, 32], 3, 2600.0, 24, 300, 1500
    n = shape[0] * shape[1]
    means = {}
    for mode in (cm.MODE_SERIAL_REFERENCE, cm.MODE_CHECKERBOARD):
        lat = make(cm, shape, V3, T, MU3, n_chains=M)
        for c in range(M):
            lat.kstate_upload(np.full(n, c % K, dtype=np.int32), c)
            lat.seed_mt19937_64(100 + c, chain=c)
        lat.seed_philox(99)
        lat.kstate_run_passes(n_eq, mode, 0)
        lat.kstate_run_passes(n_meas, mode, 1)
        per_chain = []
        for c in range(M):
            counts, bonds = lat.kstate_samples(c)
            assert counts.shape == (n_meas, K)
            phi = np.array([oracle.kstate_potential(2, K, V3, MU3, counts[i], bonds[i]) for i in range(0, n_meas, 50)]) / n
            per_chain.append(np.concatenate([counts.mean(axis=0) / n, [phi.mean()]]))
        means[mode] = np.array(per_chain)
        lat.close()
    a, b = means[cm.MODE_SERIAL_REFERENCE], means[cm.MODE_CHECKERBOARD]
    diff = a.mean(axis=0) - b.mean(axis=0)
    sigma = np.sqrt(a.var(axis=0, ddof=1) / M + b.var(axis=0, ddof=1) / M)
    assert np.all(np.abs(diff) < 3.0 * sigma), (diff, sigma)
    assert np.all(sigma[:K] < 0.02)  # the comparison is tight enough to mean something


def test_kstate_errors(cm):
    lat = cm.IsingLatticeGPU([8, 8])
    with pytest.raises(cm.CmgError):
        lat.kstate_run_passes(1)  # model not set
    with pytest.raises(cm.CmgError):
        lat.kstate_set_model(np.array([[0.0, 1.0], [2.0, 0.0]]))  # not symmetric
    lat.kstate_set_model(V3)
    with pytest.raises(cm.CmgError):
        lat.kstate_run_passes(1)  # conditions not set
    lat.kstate_set_conditions(2000.0, MU3)
    with pytest.raises(cm.CmgError):
        lat.kstate_upload(np.full(64, 3, dtype=np.int32))  # species index out of range
    with pytest.raises(cm.CmgError):
        lat.kstate_run_passes(1, cm.MODE_SERIAL_REFERENCE)  # engine not seeded
