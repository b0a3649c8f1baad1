"""events.Conversions, host evaluation (no GPU needed).

Known answers are the reference's own (python/tests/events/test_Conversions.py:20-169):
l_size, l = b * n_unitcells + unit-cell index, periodic wrap, asym_size for equal /
re-ordered occupant lists.  libcasm.xtal is absent, so the prim is given as arrays.
The rest pins the restated index arithmetic (include/casm_monte_b200/snf.hh) by
its defining properties for general integer transformation matrices: bijection,
invariance under supercell translations, unit cells inside the supercell.
"""
import itertools

import numpy as np
import pytest


@pytest.fixture(scope="module")
def events():
    import casmcode_monte_b200.monte.events as ev

    return ev


T333 = np.diag([3, 3, 3])


def test_constructor_1(events):
    # test_Conversions.py:8-20: simple cubic, occ_dof ["A", "B"]
    convert = events.Conversions(occ_dof=[["A", "B"]], transformation_matrix_to_super=T333)
    assert convert.l_size() == 27
    assert convert.asym_size() == 1 and convert.species_list() == ["A", "B"]


def test_constructor_2(events):
    # test_Conversions.py:23-72
    convert = events.Conversions(
        occ_dof=[["A", "B"], ["B", "C"]],
        transformation_matrix_to_super=T333,
        lattice_column_vector_matrix=np.eye(3),
        coordinate_frac=np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]]).T,
    )
    assert convert.l_size() == 54
    assert convert.bijk_to_l([1, 0, 0, 0]) == 27
    assert convert.l_to_bijk(27) == [1, 0, 0, 0]
    assert convert.bijk_to_l([1, 1, 0, 0]) == convert.bijk_to_l(list(np.array([1, 0, 0, 0]) + np.array([0, 1, 0, 0])))
    # periodic wrap
    assert convert.bijk_to_l([1, 3, 0, 0]) == convert.bijk_to_l([1, 0, 0, 0])
    assert convert.bijk_to_l([0, -1, 4, -7]) == convert.bijk_to_l([0, 2, 1, 2])
    # diag(n, n, n) is in Smith normal form: first index fastest
    assert convert.l_to_ijk(1) == [1, 0, 0] and convert.l_to_ijk(3) == [0, 1, 0] and convert.l_to_ijk(9) == [0, 0, 1]
    assert convert.l_to_b(30) == 1
    # coordinates
    assert np.allclose(convert.l_to_frac(27 + 1), [1.5, 0.5, 0.5])
    assert np.allclose(convert.l_to_cart(27 + 3), [0.5, 1.5, 0.5])
    assert np.allclose(convert.l_to_basis_frac(40), [0.5, 0.5, 0.5])
    # species / occupation tables (Conversions.cc:147-172)
    assert convert.species_list() == ["A", "B", "C"] and convert.species_size() == 3
    assert convert.asym_size() == 2
    a0, a1 = convert.l_to_asym(0), convert.l_to_asym(27)
    assert {a0, a1} == {0, 1}
    assert [convert.occ_to_species_index(a0, o) for o in range(convert.occ_size(a0))] == [0, 1]
    assert [convert.occ_to_species_index(a1, o) for o in range(convert.occ_size(a1))] == [1, 2]
    assert convert.species_to_occ_index(a1, 2) == 1 and convert.species_to_occ_index(a1, 0) == convert.occ_size(a1)
    assert convert.species_allowed(a0, 0) and not convert.species_allowed(a0, 2)
    assert convert.species_name_to_index("C") == 2 and convert.species_index_to_name(1) == "B"
    assert convert.species_index_to_atoms_size(0) == 1
    assert convert.asym_to_b(a1) == {1} and convert.asym_to_unitl(a0) == {0}
    assert convert.unitl_size() == 2 and convert.unitl_to_b(1) == 1 and convert.unitl_to_bijk(1) == [1, 0, 0, 0]
    assert convert.l_to_unitl(27 + 5) == 1 and convert.bijk_to_unitl([0, 7, -2, 3]) == 0
    assert convert.bijk_to_asym([1, 2, 2, 2]) == a1


def test_constructor_3_asym_by_occupant_order(events):
    # test_Conversions.py:75-169
    frac = np.array([[0.0, 0.0, 0.0], [0.0, 0.5, 0.5], [0.5, 0.0, 0.5], [0.5, 0.5, 0.0]]).T
    mol = ["A", "mol.x", "mol.y", "mol.z"]
    convert = events.Conversions(occ_dof=[["A"], mol, mol, mol], transformation_matrix_to_super=T333, coordinate_frac=frac)
    assert convert.l_size() == 27 * 4 and convert.asym_size() == 2
    convert = events.Conversions(
        occ_dof=[["A"], mol, ["mol.x", "mol.y", "mol.z", "A"], mol], transformation_matrix_to_super=T333, coordinate_frac=frac
    )
    assert convert.l_size() == 27 * 4 and convert.asym_size() == 3
    # make_with_custom_asym: reduced symmetry given explicitly
    convert = events.Conversions.make_with_custom_asym(
        occ_dof=[["A"], mol, mol, mol], transformation_matrix_to_super=T333, b_to_asym=[0, 1, 2, 1]
    )
    assert convert.asym_size() == 3 and convert.asym_to_b(1) == {1, 3}


@pytest.mark.parametrize(
    "T",
    [
        np.array([[-1, 1, 1], [1, -1, 1], [1, 1, -1]]),  # fcc primitive -> conventional, det 4
        np.array([[2, 0, 0], [0, 3, 0], [0, 0, 1]]),  # diagonal, not in Smith normal form
        np.array([[2, 1, 0], [0, 3, 1], [1, 0, 2]]),  # det 13
        np.array([[0, 2, 0], [-3, 0, 0], [0, 0, 4]]),  # negative determinant... (det 24)
        np.array([[4, 0, 0], [0, 4, 0], [0, 0, 2]]),
    ],
)
def test_general_transformation_matrices(events, T):
    nb = 2
    convert = events.Conversions(occ_dof=[["A", "B"]] * nb, transformation_matrix_to_super=T)
    n_uc = abs(round(np.linalg.det(T)))
    assert convert.l_size() == nb * n_uc
    seen = set()
    Tinv = np.linalg.inv(T)
    for l in range(convert.l_size()):
        b, i, j, k = convert.l_to_bijk(l)
        assert b == l // n_uc  # l = b * n_unitcells + unit-cell index
        assert convert.bijk_to_l([b, i, j, k]) == l
        frac = Tinv @ np.array([i, j, k])
        assert np.all(frac > -1e-9) and np.all(frac < 1 - 1e-9)  # the unit cell lies inside the supercell
        seen.add((b, i, j, k))
    assert len(seen) == convert.l_size()
    # invariance under supercell lattice translations T * n
    for n in itertools.product((-2, 0, 1), repeat=3):
        shift = T @ np.array(n)
        for l in (0, convert.l_size() // 2, convert.l_size() - 1):
            b, i, j, k = convert.l_to_bijk(l)
            assert convert.bijk_to_l([b, i + shift[0], j + shift[1], k + shift[2]]) == l


def test_custom_unitcell(events):
    # a 2 x 1 x 1 unit supercell inside a 4 x 2 x 2 supercell: alternating orbits along a
    T = np.diag([4, 2, 2])
    U = np.diag([2, 1, 1])
    convert = events.Conversions.make_with_custom_unitcell(
        occ_dof=[["A", "B"]], species_list=["A", "B"], transformation_matrix_to_super=T,
        unit_transformation_matrix_to_super=U, unitl_to_asym=[0, 1],
    )
    assert convert.unitl_size() == 2 and convert.asym_size() == 2
    for l in range(convert.l_size()):
        b, i, j, k = convert.l_to_bijk(l)
        assert convert.l_to_asym(l) == i % 2
    assert np.array_equal(convert.unit_transformation_matrix_to_super(), U)
    with pytest.raises(RuntimeError):
        events.Conversions.make_with_custom_unitcell(
            occ_dof=[["A", "B"]], species_list=["A", "B"], transformation_matrix_to_super=np.diag([3, 2, 2]),
            unit_transformation_matrix_to_super=U, unitl_to_asym=[0, 1],
        )  # U does not tile S


def test_errors(events):
    with pytest.raises(RuntimeError):
        events.Conversions(occ_dof=[["A"]], transformation_matrix_to_super=np.zeros((3, 3)))
    c = events.Conversions(occ_dof=[["A", "B"]], transformation_matrix_to_super=T333)
    with pytest.raises(RuntimeError):
        c.l_to_bijk(27)
    with pytest.raises(RuntimeError):
        c.bijk_to_l([1, 0, 0, 0])
    # round-1 constructor still there
    f = events.Conversions([3, 3, 3], n_basis=2)
    assert f.l_size() == 54 and f.bijk_to_l([1, 0, 0, 0]) == 27
