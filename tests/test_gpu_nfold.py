"""N-fold way (rejection-free) driver on the device (cmg_nfold_run) against the CPU restatement
of the reference's loop (include/casm/monte/methods/nfold.hh:80-147; oracle/kstate_oracle.hh,
namespace nfold): the sequence of events, the sampled integer observables, the final occupation
and the engine state are identical; the sample weights (time increments) agree to the last bits
of log(); and the weighted averages agree with Metropolis sampling within 3 sigma, as does the
expected acceptance rate with the measured Metropolis acceptance rate."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

J = 0.1


@pytest.fixture(scope="module")
def cm():
    import casmcode_monte_b200 as m

    return m


@pytest.mark.parametrize("shape,T,mu", [([16, 12], 2000.0, 0.0), ([9, 7], 2633.0, 0.05), ([6, 6, 4], 5235.0, -0.02), ([32, 32], 1200.0, 0.0)])
def test_nfold_reproduces_the_restated_driver(cm, oracle, shape, T, mu):
    n = int(np.prod(shape))
    occ = np.random.default_rng(n).choice(np.array([-1, 1], dtype=np.int32), size=n)
    lat = cm.IsingLatticeGPU(shape, J=J)
    lat.set_conditions(T, mu)
    lat.seed_mt19937_64(99)
    lat.upload(occ)
    # two calls: class lists, clock and sample schedule continue
    lat.nfold_run(700, 7)
    lat.nfold_run(1301, 7)
    e = oracle.RandomNumberEngine()
    e.seed(99)
    ref = oracle.nfold_run(shape, occ, J, T, mu, e, 2001, 7, True)
    assert np.array_equal(lat.download(), ref["occupation"])
    S, B = lat.samples_sb()
    assert len(S) == 2001 // 7 and np.array_equal(S, ref["S"]) and np.array_equal(B, ref["B"])
    w, r = lat.nfold_weights()
    assert np.allclose(w, ref["weight"], rtol=1e-14, atol=0.0) and np.all(w > 0)
    assert np.allclose(r, ref["expected_acceptance_rate"], rtol=1e-15, atol=0.0)
    t, steps = lat.nfold_time()
    assert steps == 2001 and np.isclose(t, ref["time"], rtol=1e-12)
    st, pos = lat.get_engine_state()
    assert e.dump().split()[:312] == [str(int(v)) for v in st] and int(e.dump().split()[312]) == pos
    # the doubles of the three default observables are available for the (weighted) statistics
    x = lat.samples(cm.Q_PARAM_COMPOSITION)
    assert np.array_equal(x, (n + S.astype(np.float64)) / 2.0 / n)


def test_nfold_weighted_averages_match_metropolis_within_3_sigma(cm):
    """Low temperature, where Metropolis rejects ~97 % of its attempts: the rejection-free driver's
    weighted averages (through the device weighted-statistics entry point) against checkerboard
    Metropolis, 16 chains each, chain-to-chain scatter as the error bar."""
    from casmcode_monte_b200.lattice import host_series_stats_weighted

    shape, T, mu, M = [32, 32], 2000.0, 0.02, 16
    n = shape[0] * shape[1]
    nf = cm.IsingLatticeGPU(shape, n_chains=M, J=J)
    nf.set_conditions(T, mu)
    for c in range(M):
        nf.seed_mt19937_64(500 + c, chain=c)
    nf.nfold_run(20000, 0)  # equilibrate: 20000 accepted events per chain
    nf.nfold_run(200000, 20)
    x_nf, e_nf, acc_nf = [], [], []
    for c in range(M):
        w, r = nf.nfold_weights(c)
        x = nf.samples(cm.Q_PARAM_COMPOSITION, c)
        e = nf.samples(cm.Q_POTENTIAL_ENERGY, c)
        x_nf.append(float((x * w).sum() / w.sum()))
        e_nf.append(float((e * w).sum() / w.sum()))
        acc_nf.append(float((r * w).sum() / w.sum()))
    st = host_series_stats_weighted(nf.samples(cm.Q_PARAM_COMPOSITION, 0), nf.nfold_weights(0)[0], method=1, n_resamples=2000)
    assert abs(st["mean"] - x_nf[0]) < 1e-12 and st["calculated_precision"] > 0
    mp = cm.IsingLatticeGPU(shape, n_chains=M, J=J)
    mp.set_conditions(T, mu)
    mp.seed_philox(123)
    mp.run_passes(2000, cm.MODE_CHECKERBOARD, 0)
    mp.reset_counters()
    mp.run_passes(20000, cm.MODE_CHECKERBOARD, 10)
    x_mp = [float(mp.samples(cm.Q_PARAM_COMPOSITION, c).mean()) for c in range(M)]
    e_mp = [float(mp.samples(cm.Q_POTENTIAL_ENERGY, c).mean()) for c in range(M)]
    acc_mp = [mp.counters(c)[1] / (mp.counters(c)[1] + mp.counters(c)[2]) for c in range(M)]
    for a, b, name in ((x_nf, x_mp, "param_composition"), (e_nf, e_mp, "potential_energy"), (acc_nf, acc_mp, "acceptance rate")):
        a, b = np.array(a), np.array(b)
        sigma = np.sqrt(a.var(ddof=1) / M + b.var(ddof=1) / M)
        assert abs(a.mean() - b.mean()) < 3.0 * sigma, (name, a.mean(), b.mean(), sigma)
    assert 0.02 < np.mean(acc_mp) < 0.06  # the regime the N-fold way is for
