"""GPU micro-sweep: throughput of the half-sweep kernel variants / strip lengths.
Usage (on the GPU box): python tools/sweep_variants.py > gpurun_out/sweep.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from casmcode_monte_b200 import MODE_CHECKERBOARD, IsingLatticeGPU


def time_passes(lat, stream, n_passes, sample_period):
    lat.run_passes(10, MODE_CHECKERBOARD, sample_period)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    lat.run_passes(n_passes, MODE_CHECKERBOARD, sample_period)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3


def main():
    stream = torch.cuda.Stream()
    out = []
    cases = [
        ("2d", [4096, 4096], 1, ["bulk2d", "bulk2d:js=8", "tile2d:p=3:nt=512"], 120),
        ("2d_sweep8", [4096, 4096], 8, ["bulk2d", "bulk2d:ns=1", "bulk2d:js=32", "tile2d:p=3:nt=512"], 48),
        ("2d_sweep8_js", [4096, 4096], 8, ["bulk2d:js=28", "bulk2d:js=56", "bulk2d:js=52", "bulk2d:js=60", "bulk2d:js=19", "bulk2d:js=14"], 48),
        ("2d_big_js", [16384, 16384], 1, ["bulk2d", "bulk2d:js=56", "bulk2d:js=112", "bulk2d:js=20"], 20),
        ("2d_grid", [256, 256], 128, ["tile2d:nt=512"], 128),
        ("2d_big", [16384, 16384], 1, ["bulk2d", "bulk2d:ns=296", "bulk2d:ns=444", "bulk2d:ns=592"], 20),
        ("3d", [512, 512, 512], 1, ["bulk3d", "bulk3d:js=8", "bulk3d:js=16"], 10),
        ("2d_slab8", [65536, 8192], 1, ["bulk2d", "bulk2d:ns=1", "bulk2d:ns=74", "bulk2d:ns=111"], 6),
        ("2d_slab2", [65536, 32768], 1, ["bulk2d", "bulk2d:ns=74", "bulk2d:ns=111", "bulk2d:ns=148", "bulk2d:ns=296"], 4),
        ("2d_slab4", [65536, 16384], 1, ["bulk2d", "bulk2d:ns=74", "bulk2d:ns=111", "bulk2d:ns=148", "bulk2d:ns=222"], 4),
        ("2d_one", [65536, 65536], 1, ["bulk2d", "bulk2d:ns=74", "bulk2d:ns=148", "bulk2d:ns=296"], 2),
        ("3d_tma", [512, 512, 512], 1, ["tma3d", "bulk3d", "tma3d:js=32", "tma3d:js=64", "tma3d:js=128"], 10),
        ("2d_tile", [4096, 4096], 1, ["tile2d:p=2:nt=512", "tile2d:p=3:nt=512", "tile2d:p=4:nt=512", "tile2d:p=3:nt=640", "tile2d:p=4:nt=640", "tile2d:p=3:nt=768", "tile2d:p=3:nt=1024", "tile2d:p=3:nt=256"], 120),
        ("3d_js", [512, 512, 512], 1, ["bulk3d:js=%d" % j for j in (24, 36, 40, 44, 48, 52, 56, 60, 64, 68, 72, 76, 80, 88, 96, 104, 128)], 10),
        ("2d_ring", [4096, 4096], 1, ["ring2d:rp=128", "ring2d:rp=32", "ring2d:rp=256", "tile2d:p=3:nt=512", "bulk2d"], 256),
        ("2d_ringsync", [4096, 4096], 1, ["ring2d", "ring2d"], 512),
        ("2d_ringsync_small", [1024, 4096], 1, ["ring2d"], 512),
        ("2d_ringsync_8192", [8192, 2048], 1, ["ring2d"], 256),
        ("pdl_3d", [512, 512, 512], 1, ["bulk3d", "bulk3d:chain=0", "bulk3d:pdl=0", "bulk3d", "bulk3d:chain=0"], 20),
        ("pdl_2d_sweep8", [4096, 4096], 8, ["bulk2d", "bulk2d:chain=0", "bulk2d:pdl=0", "bulk2d", "bulk2d:chain=0"], 48),
        ("pdl_2d_one", [4096, 4096], 1, ["bulk2d", "bulk2d:chain=0", "bulk2d:pdl=0"], 120),
        ("pdl_3d_256", [256, 256, 256], 1, ["bulk3d", "bulk3d:chain=0", "bulk3d:pdl=0"], 40),
        ("chain_ns_2d", [4096, 4096], 8, ["bulk2d", "bulk2d:ns=92", "bulk2d:ns=111", "bulk2d:ns=148", "bulk2d:ns=56", "bulk2d:ns=222"], 48),
        ("chain_ns_3d", [512, 512, 512], 1, ["bulk3d", "bulk3d:ns=12", "bulk3d:ns=14", "bulk3d:ns=18", "bulk3d:ns=7", "bulk3d:ns=27"], 20),
        ("chain_slab8", [65536, 8192], 1, ["bulk2d", "bulk2d:chain=0", "bulk2d:pdl=0"], 12),
        ("chain_slab4", [65536, 16384], 1, ["bulk2d", "bulk2d:chain=0", "bulk2d:pdl=0"], 8),
        ("pdl_tile_512", [512, 512], 1, ["auto", "auto:pdl=0", "tile2d:p=3", "tile2d:p=3:pdl=0", "tile2d:p=2", "tile2d:p=2:pdl=0"], 600),
        ("pdl_tile_768", [768, 2048], 1, ["auto", "auto:pdl=0"], 300),
        ("pdl_tile_grid", [256, 256], 128, ["auto", "auto:pdl=0"], 128),
        ("tile_p_512", [512, 512], 1, ["tile2d:p=3", "tile2d:p=4", "tile2d:p=5", "tile2d:p=6", "tile2d:p=7", "tile2d:p=8", "tile2d:p=10", "tile2d:p=12"], 840),
        ("tile_p_768", [768, 2048], 1, ["tile2d:p=3", "tile2d:p=4", "tile2d:p=5", "tile2d:p=6", "tile2d:p=8"], 240),
        ("tile_p_256", [256, 1024], 1, ["tile2d:p=3", "tile2d:p=5", "tile2d:p=7", "tile2d:p=10"], 840),
        ("tile_p_2048", [2048, 2048], 1, ["tile2d:p=3", "tile2d:p=4", "tile2d:p=5", "ring2d"], 240),
        ("tile_nt_512", [512, 512], 1, ["tile2d:nt=512", "tile2d:nt=256", "tile2d:nt=256:p=5"], 840),
        ("tile_nt_768", [768, 2048], 1, ["tile2d:nt=512", "tile2d:nt=256"], 240),
        ("tile_nt_256", [256, 1024], 1, ["tile2d:nt=512", "tile2d:nt=256"], 840),
        ("mid_ring", [512, 512], 1, ["auto", "ring2d", "tile2d"], 840),
        ("mid_ring_256", [256, 1024], 1, ["auto", "tile2d"], 840),
        ("mid_ring_512x2048", [512, 2048], 1, ["auto", "tile2d"], 400),
        ("mid_ring_256sq", [256, 256], 1, ["auto", "ring2d"], 840),
        ("pdl_2d_16k", [16384, 16384], 1, ["bulk2d", "bulk2d:chain=0", "bulk2d:pdl=0"], 20),
        ("2d_gridtile", [256, 256], 128, ["tile2d:nt=256", "tile2d:nt=512", "tile2d:nt=640", "tile2d:nt=768", "tile2d:nt=1024"], 128),
    ]
    only = sys.argv[1].split(",") if len(sys.argv) > 1 else None  # e.g. "2d_sweep8,3d"
    auto_only = len(sys.argv) > 2 and sys.argv[2] == "auto"       # first listed variant only
    for name, shape, chains, variants, n_passes in cases:
        if only and name not in only:
            continue
        if auto_only:
            variants = variants[:1]
        lat = IsingLatticeGPU(shape, n_chains=chains, J=0.1)
        lat.set_stream(stream.cuda_stream)
        lat.set_conditions(2633.0 if len(shape) == 2 else 5235.0, 0.0)
        lat.seed_philox(1)
        lat.randomize(5, 0.5)
        n = chains
        for s in shape:
            n *= s
        for v in variants:
            lat.set_kernel_variant(v)
            for sp in (0, 1):
                t = time_passes(lat, stream, n_passes, sp)
                rec = {"case": name, "variant": v, "sample_period": sp, "attempts_per_s": n * n_passes / t, "us_per_halfsweep": t / (2 * n_passes) * 1e6}
                out.append(rec)
                print(json.dumps(rec), flush=True)
        lat.close()


if __name__ == "__main__":
    main()
