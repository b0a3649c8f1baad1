"""Short GPU run for ncu captures: a few checkerboard passes of the headline
workload (4096x4096, T=2633 K), or of the case named on the command line."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from casmcode_monte_b200 import MODE_CHECKERBOARD, IsingLatticeGPU

case = sys.argv[1] if len(sys.argv) > 1 else "2d"
n_passes = int(sys.argv[2]) if len(sys.argv) > 2 else 30
variant = sys.argv[3] if len(sys.argv) > 3 else "auto"
shape, chains, T = {"2d": ([4096, 4096], 1, 2633.0), "3d": ([512, 512, 512], 1, 5235.0), "grid": ([256, 256], 128, 2633.0), "sweep8": ([4096, 4096], 8, 2633.0), "big": ([16384, 16384], 1, 2633.0), "mid": ([512, 512], 1, 2633.0),
                    # the slab of one rank of the 65536^2 decomposed run at N = 8 / 4 / 2 GPUs (same kernel, same grid)
                    "slab8": ([65536, 8192], 1, 2633.0), "slab4": ([65536, 16384], 1, 2633.0), "slab2": ([65536, 32768], 1, 2633.0)}[case]
lat = IsingLatticeGPU(shape, n_chains=chains, J=0.1)
lat.set_conditions(T, 0.0)
lat.seed_philox(0xC0FFEE)
lat.randomize(12345, 0.5)
lat.set_kernel_variant(variant)
lat.run_passes(n_passes, MODE_CHECKERBOARD, 1)
lat.sync()
print(lat.kernel_variant, lat.counters())
