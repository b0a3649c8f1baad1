"""Host mirror of the general multi-species proposal machinery (include/casm_monte_b200/
events.hh): OccCandidateList, swaps, OccLocation, propose_*_event.  Known structure from the
reference's sources (src/casm/monte/events/OccCandidate.cc:32-182, OccLocation.cc:39-116,
:253-283; include/casm/monte/events/OccEventProposal.hh:108-348); the proposal stream is
compared draw for draw with the CPU oracle's restatement."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def ev():
    import casmcode_monte_b200.monte.events as events

    return events


@pytest.fixture(scope="module")
def monte():
    import casmcode_monte_b200.monte as m

    return m


def two_sublattice_convert(ev):
    # the prim of python/tests/events/test_Conversions.py:23-72
    return ev.Conversions(occ_dof=[["A", "B"], ["B", "C"]], transformation_matrix_to_super=np.diag([3, 3, 3]))


def test_candidates_and_swaps(ev):
    convert = two_sublattice_convert(ev)
    cands = ev.OccCandidateList(convert)
    assert [c.to_tuple() for c in cands] == [(0, 0), (0, 1), (1, 1), (1, 2)]  # (asym, species): A, B | B, C
    assert len(cands) == 4 and cands.index(ev.OccCandidate(1, 2)) == 3
    assert cands.matching_index(0, 2) == len(cands)  # C is not allowed on the first orbit
    assert ev.OccCandidate(1, 0).is_valid(convert) is False and ev.OccCandidate(0, 1).is_valid(convert)
    can = ev.make_canonical_swaps(convert, cands)
    assert [s.to_tuple() for s in can] == [(0, 0, 0, 1), (1, 1, 1, 2)]  # a < b only, species allowed on both sites
    sgc = ev.make_semigrand_canonical_swaps(convert, cands)
    assert [s.to_tuple() for s in sgc] == [(0, 0, 0, 1), (0, 1, 0, 0), (1, 1, 1, 2), (1, 2, 1, 1)]  # both directions
    assert ev.get_n_allowed_per_unitcell(convert, sgc) == 2
    assert not ev.is_allowed_canonical_swap(convert, ev.OccCandidate(0, 0), ev.OccCandidate(1, 1))  # A not allowed on orbit 1
    assert not ev.is_allowed_semigrand_canonical_swap(convert, ev.OccCandidate(0, 1), ev.OccCandidate(1, 2))  # different orbits
    s = ev.OccSwap(ev.OccCandidate(1, 2), ev.OccCandidate(0, 1))
    assert s.sorted().to_tuple() == (0, 1, 1, 2) and s.to_tuple() == (1, 2, 0, 1)
    # custom candidate list
    custom = ev.OccCandidateList([ev.OccCandidate(1, 2), ev.OccCandidate(0, 0)], convert)
    assert custom.index(ev.OccCandidate(0, 0)) == 1 and custom.index(ev.OccCandidate(0, 1)) == 2


def check_location_invariants(ev, convert, cands, loc, occ):
    n_mut = loc.mol_size()
    seen = 0
    for ci in range(len(cands)):
        cand = cands[ci]
        for pos in range(loc.cand_size(cand)):
            mol = loc.mol(loc.mol_id(cand, pos))
            assert mol.mol_location_index == pos and mol.asymmetric_unit_index == cand.asymmetric_unit_index
            assert mol.species_index == cand.species_index
            l = mol.linear_site_index
            assert convert.occ_to_species_index(convert.l_to_asym(l), int(occ[l])) == mol.species_index
            assert loc.linear_site_index_to_mol_id(l) == mol.id
            seen += 1
    assert seen == n_mut


def test_occ_location_tracks_semigrand_and_canonical_events(ev, monte):
    convert = two_sublattice_convert(ev)
    cands = ev.OccCandidateList(convert)
    rng_np = np.random.default_rng(3)
    occ = rng_np.integers(0, 2, size=convert.l_size()).astype(np.int32)
    loc = ev.OccLocation(convert, cands)
    loc.initialize(occ)
    assert loc.mol_size() == 54
    assert sum(loc.cand_size(c) for c in cands) == 54
    assert loc.cand_size(ev.OccCandidate(0, 0)) == int((occ[:27] == 0).sum())
    check_location_invariants(ev, convert, cands, loc, occ)
    engine = monte.RandomNumberEngine()
    engine.seed(11)
    rng = monte.RandomNumberGenerator(engine)
    sgc = ev.make_semigrand_canonical_swaps(convert, cands)
    can = ev.make_canonical_swaps(convert, cands)
    e = ev.OccEvent()
    for it in range(300):
        n_before = [loc.cand_size(c) for c in cands]
        if it % 2 == 0:
            ev.propose_semigrand_canonical_event(e, loc, sgc, rng)
            assert len(e.linear_site_index) == 1 and len(e.occ_transform) == 1
            t = e.occ_transform[0]
            assert t.linear_site_index == e.linear_site_index[0] and t.from_species != t.to_species
            assert convert.species_to_occ_index(t.asym, t.to_species) == e.new_occ[0]
            assert convert.occ_to_species_index(t.asym, int(occ[t.linear_site_index])) == t.from_species
        else:
            ev.propose_canonical_event(e, loc, can, rng)
            assert len(e.linear_site_index) == 2
            a, b = e.occ_transform
            assert (a.from_species, a.to_species) == (b.to_species, b.from_species)
        loc.apply(e, occ)
        for l, o in zip(e.linear_site_index, e.new_occ):
            assert occ[l] == o
        n_after = [loc.cand_size(c) for c in cands]
        assert sum(n_after) == 54 and (it % 2 == 0 or n_after == n_before)  # canonical events conserve the counts
    check_location_invariants(ev, convert, cands, loc, occ)
    with pytest.raises(RuntimeError):
        ev.OccLocation(convert, cands, update_atoms=True)  # atom trajectories (KMC) are out of scope
    bad = ev.OccEvent()
    ev.propose_semigrand_canonical_event(bad, loc, sgc, rng)
    tr = bad.occ_transform
    tr[0].from_species = 2 if tr[0].from_species != 2 else 0
    bad.occ_transform = tr
    with pytest.raises(RuntimeError):
        loc.apply(bad, occ)  # OccLocation.cc:259-261 species mismatch


@pytest.mark.parametrize("K", [2, 3, 4])
def test_proposal_stream_equals_the_restated_reference(ev, monte, oracle, K):
    """Same engine state => the same sequence of proposed events as the oracle's restated
    choose_semigrand_canonical_swap / choose_mol / apply (draw for draw)."""
    names = ["A", "B", "C", "D"][:K]
    convert = ev.Conversions(occ_dof=[names], transformation_matrix_to_super=np.diag([4, 3, 2]))
    cands = ev.OccCandidateList(convert)
    swaps = ev.make_semigrand_canonical_swaps(convert, cands)
    assert [s.to_tuple() for s in swaps] == [tuple(t) for t in oracle.kstate_swaps(K)]
    occ = np.random.default_rng(K).integers(0, K, size=24).astype(np.int32)
    loc = ev.OccLocation(convert, cands)
    loc.initialize(occ)
    engine = monte.RandomNumberEngine()
    engine.seed(2024)
    rng = monte.RandomNumberGenerator(engine)
    oe = oracle.RandomNumberEngine()
    oe.seed(2024)
    ref_events, ref_occ = oracle.kstate_propose_sequence(K, occ, oe, 500)
    e = ev.OccEvent()
    mine = occ.copy()
    got = []
    for _ in range(500):
        ev.propose_semigrand_canonical_event(e, loc, swaps, rng)
        got.append((e.linear_site_index[0], e.new_occ[0]))
        loc.apply(e, mine)
    assert got == [tuple(t) for t in ref_events]
    assert np.array_equal(mine, ref_occ)
    assert engine.dump() == oe.dump()
