"""Whatever kernel and launch form `auto` selects for a shape and a number of chains (resident
ring, one lattice per CTA, tiles with halos, chained streaming launches, plain ones), and every
forced 2-d variant that accepts the shape, must leave the occupation, the sampled sums and the
counters of the generic byte kernel -- which the other tests pin to the oracle.  The tiled
kernel with halos is the reason this file exists: it wrote its columns back in place, which is
only right while every tile of a lattice stages its halos before any tile finishes; 16 chains
of 1024^2 (2368 CTAs, 16 waves) differed from run to run until the write-back went to a
second copy of the planes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [([2048, 2048], 1, 12), ([8192, 4096], 1, 5), ([1024, 1024], 16, 12), ([512, 512, 64], 1, 8), ([128, 128, 128], 2, 10),
         ([64, 64], 512, 20), ([100, 100], 3, 8), ([4096, 64], 1, 12), ([512, 512], 1, 14), ([768, 2048], 2, 9),
         ([256, 256], 300, 12), ([96, 34], 5, 8), ([2048, 64, 16], 1, 8), ([32, 6], 1, 9)]


def run(cm, shape, chains, n_passes, variant):
    lat = cm.IsingLatticeGPU(shape, n_chains=chains, J=0.1)
    for ch in range(chains):
        lat.set_conditions((2633.0 if len(shape) == 2 else 5235.0) + 7.0 * ch, 0.002 * (ch % 5), chain=ch)
        lat.randomize(11 + ch, 0.5, chain=ch)
    lat.seed_philox(2024)
    lat.set_kernel_variant(variant)
    used = []
    for n, sp in ((n_passes, 2), (3, 1)):  # a long call, then one shorter than the resident kernel takes
        lat.run_passes(n, cm.MODE_CHECKERBOARD, sp)
        used.append(lat.kernel_variant)
    lat.sync()
    pick = sorted({0, chains // 2, chains - 1})
    out = ([lat.download(ch) for ch in pick], [lat.samples_sb(ch) for ch in pick], [lat.counters(ch) for ch in pick], used)
    lat.close()
    return out


@pytest.mark.parametrize("shape,chains,n_passes", CASES)
def test_auto_selection_equals_the_generic_kernel(shape, chains, n_passes):
    import casmcode_monte_b200 as cm

    ref = run(cm, shape, chains, n_passes, "generic")
    variants = ["auto"]
    if len(shape) == 2 and shape[0] % 64 == 0:
        variants += ["tile2d", "tile2d:pdl=0"]
    if len(shape) == 2 and shape[0] % 32 == 0:
        variants += ["bulk2d"]
    for v in variants:
        try:
            got = run(cm, shape, chains, n_passes, v)
        except RuntimeError as e:  # a forced variant that does not fit the shape refuses it
            assert v != "auto", e
            continue
        for a, b in zip(got[0], ref[0]):
            assert np.array_equal(a, b), (v, got[3])
        for a, b in zip(got[1], ref[1]):
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (v, got[3])
        assert got[2] == ref[2], (v, got[3])
