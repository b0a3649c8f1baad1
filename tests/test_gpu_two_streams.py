"""Two contexts on two streams share the GPU -- chained streaming launches next to tiles with
halos, next to the resident kernel, next to themselves: each ends exactly where it ends when it
runs alone (no wait of one context can be satisfied or starved by the other's CTAs)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make(cm, shape, chains, variant, stream, seed):
    lat = cm.IsingLatticeGPU(shape, n_chains=chains, J=0.1)
    lat.set_stream(stream.cuda_stream)
    for ch in range(chains):
        lat.set_conditions((2633.0 if len(shape) == 2 else 5235.0) + 5.0 * ch, 0.001 * ch, chain=ch)
        lat.randomize(seed + ch, 0.5, chain=ch)
    lat.seed_philox(seed)
    lat.set_kernel_variant(variant)
    return lat


def result(lat, chains):
    lat.sync()
    return [lat.download(ch) for ch in range(chains)], [lat.samples_sb(ch) for ch in range(chains)]


@pytest.mark.parametrize("a,b", [(([4096, 2048], 4, "bulk2d"), ([512, 256, 64], 1, "bulk3d")), (([1024, 1024], 8, "tile2d"), ([4096, 1024], 2, "bulk2d")),
                                 (([4096, 1024], 1, "ring2d"), ([1024, 512], 4, "tile2d")), (([2048, 2048], 4, "bulk2d"), ([2048, 2048], 4, "bulk2d"))])
def test_two_contexts_on_two_streams_end_where_they_end_alone(a, b):
    import torch

    import casmcode_monte_b200 as cm

    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    alone = []
    for (shape, chains, variant), stream, seed in ((a, s1, 5), (b, s2, 9)):
        lat = make(cm, shape, chains, variant, stream, seed)
        for _ in range(5):
            lat.run_passes(8, cm.MODE_CHECKERBOARD, 2)
        alone.append(result(lat, chains))
        lat.close()
    la, lb = make(cm, *a, s1, 5), make(cm, *b, s2, 9)
    for _ in range(5):  # interleaved enqueues: the kernels of the two streams share the GPU
        la.run_passes(8, cm.MODE_CHECKERBOARD, 2)
        lb.run_passes(8, cm.MODE_CHECKERBOARD, 2)
    together = [result(la, a[1]), result(lb, b[1])]
    la.close()
    lb.close()
    for (o1, sb1), (o2, sb2) in zip(alone, together):
        assert all(np.array_equal(x, y) for x, y in zip(o1, o2))
        assert all(np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) for x, y in zip(sb1, sb2))
