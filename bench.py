#!/usr/bin/env python
"""bench.py -- Metropolis flip attempts/s of the Ising SGC hot path on B200.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W`
prints ONE JSON line.  A "step" is PASSES_PER_STEP checkerboard passes (one
pass = one attempted flip per site, on-device sampling of energy and
composition every pass) over one synthetic lattice.

Workload (BASELINE.json configs[1]): the 2-d square 4096x4096 SGC Ising
temperature sweep through T_c, J = 0.1 eV: the eight temperatures of the sweep
(1800 ... 3200 K, T_c = 2633 K among them) are eight independent lattices on
one GPU, mu = 0, swept one after the other; the lattice being swept lives in the
shared memory of the whole GPU (k_ring2d, a cooperative launch of up to 128
passes).  With N > 1 GPUs every rank runs the same eight-temperature sweep at
its own exchange potential mu_r (a (T, mu) phase-diagram grid sharded over
GPUs, no communication: weak scaling).  Reported next to it: the same eight
lattices advanced concurrently by the HBM-streaming strip kernel
(`streaming_8_lattices`, k_halfsweep_bulk2d, with its DRAM traffic) and a
single lattice at T_c alone (`single_lattice`).

`--impl reference` times the CPU restatement of the reference loop
(oracle/oracle_bench, one independent chain per host core) on a bounded sample
of the same workload; it is the reference arm, not the product.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N0 = N1 = 4096
J = 0.1
MU = 0.0
T_HEADLINE = 2633.0
T_SWEEP = [1800.0, 2200.0, 2500.0, 2600.0, 2633.0, 2660.0, 2800.0, 3200.0]
MU_GRID = [0.0, 0.02, -0.02, 0.04, -0.04, 0.06, -0.06, 0.08]
PASSES_PER_STEP = 500
ALGO_BYTES_PER_ATTEMPT = 3.0  # int8, two colour planes: read own + read other + write own
METRIC = "metropolis_flip_attempts_per_s"
UNIT = "attempts/s"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons while the timed region runs: NVML in
    process every 10 ms (the timed region is well under a second), nvidia-smi as
    the fallback."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self._stop_evt = threading.Event()
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        sm = float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
        mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
        bits = [
            getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        ]
        self.rows.append([str(sm), str(self._max)] + ["Active" if mask & b else "Not Active" for b in bits])

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                    self._stop_evt.wait(0.01)
                    continue
                out = subprocess.run(
                    ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                    capture_output=True, text=True, timeout=5,
                ).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                self._nvml = None
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for nm, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_cpu_reference(n0, n1, T, mu, n_passes, sample_period, use_nlist, threads, seed=12345):
    exe = os.path.join(ROOT, "oracle", "oracle_bench")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j2"], stdout=subprocess.DEVNULL)
    out = subprocess.run(
        [exe, str(n0), str(n1), str(T), str(mu), str(n_passes), str(sample_period), str(int(use_nlist)), str(threads), str(seed)],
        capture_output=True, text=True, check=True,
    ).stdout
    return json.loads(out.strip().splitlines()[-1])


def reference_arm(args):
    """CPU implementation of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    total_steps = args.steps + args.warmup
    if total_steps <= 12:
        n0, n1, passes = N0, N1, 1
        sample = f"{cores} independent chains (one per core) x 1 pass of the full {N0}x{N1} lattice per step, use_nlist=false, sampling every pass"
    else:
        n0, n1, passes = 1024, 1024, 2
        sample = f"{cores} independent chains x 2 passes of a 1024x1024 proxy lattice per step (full size would exceed the time budget), use_nlist=false, sampling every pass"
    for _ in range(args.warmup):
        run_cpu_reference(n0, n1, T_HEADLINE, MU, passes, 1, False, cores)
    t_total, attempts = 0.0, 0.0
    for i in range(args.steps):
        r = run_cpu_reference(n0, n1, T_HEADLINE, MU, passes, 1, False, cores, seed=1000 + i)
        t_total += r["wall_s"]
        attempts += float(n0) * n1 * passes * cores
    value = attempts / t_total
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / max(args.steps, 1),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "int32 occupation / f64 energies (CPU)",
        "data": "synthetic",
        "config": workload_config(1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {
        "workload": f"2D square Ising SGC temperature sweep through T_c, {N0}x{N1} supercell, checkerboard sweeps: "
        f"{len(T_SWEEP)} temperatures {T_SWEEP} K as {len(T_SWEEP)} lattices per GPU swept one after the other "
        f"(the lattice being swept is resident in shared memory), J=0.1 eV, "
        + ("mu=0" if n_gpus == 1 else f"rank r at mu_r={MU_GRID[:n_gpus]} eV ((T, mu) grid sharded over {n_gpus} GPUs, no communication)"),
        "lattice": [N0, N1],
        "lattices_per_gpu": len(T_SWEEP),
        "passes_per_step": PASSES_PER_STEP,
        "sample_period": 1,
        "initial_state": "i.i.d. +1/-1 (Philox, seed 12345)",
        "philox_seed": "0xC0FFEE + rank",
        "l2": "L2 flushed (256 MiB write) between timed steps; the 8 x 16 MiB int8 planes (128 MiB) slightly exceed the 126 MB L2, "
              "every launch stages its lattice from global memory",
        "timing": "CUDA events on the launching stream per step, summed; max over ranks",
    }


def ours(args):
    import numpy as np
    import torch

    from casmcode_monte_b200 import MODE_CHECKERBOARD, IsingLatticeGPU

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    mu = MU if world == 1 else MU_GRID[rank % len(MU_GRID)]
    n_lat = len(T_SWEEP)
    stream = torch.cuda.Stream()
    lat = IsingLatticeGPU([N0, N1], n_chains=n_lat, device=local_rank, J=J)
    lat.set_stream(stream.cuda_stream)
    for ch, T in enumerate(T_SWEEP):
        lat.set_conditions(T, mu, chain=ch)
    lat.seed_philox(0xC0FFEE + rank)
    lat.randomize(12345 + rank, 0.5)
    lat.sync()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    n_sites = N0 * N1 * n_lat  # sites per pass over all lattices of this GPU

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # the eight lattices as eight one-lattice contexts on the timing stream (same
    # global Philox streams as chains 0..7 of `lat` through set_chain_offset)
    per_lat = N0 * N1
    host_occ = torch.empty((n_lat, per_lat), dtype=torch.int32).pin_memory()
    sweep = []
    for k, T in enumerate(T_SWEEP):
        lat.download(k, out=host_occ.numpy()[k])
        lk = IsingLatticeGPU([N0, N1], device=local_rank, J=J)
        lk.set_stream(stream.cuda_stream)
        lk.set_conditions(T, mu)
        lk.seed_philox(0xC0FFEE + rank)
        lk.set_chain_offset(k)
        if args.value_variant != "auto":
            lk.set_kernel_variant(args.value_variant)
        lk.upload(host_occ.numpy()[k])
        sweep.append(lk)

    def one_step(ctxs):
        with torch.cuda.stream(stream):
            flush.fill_(1)  # L2 flush, outside the events
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for c in ctxs:
                c.run_passes(PASSES_PER_STEP, MODE_CHECKERBOARD, 1)
            e1.record(stream)
        return (e0, e1)

    for _ in range(args.warmup):
        one_step(sweep)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = sum(c.launch_count for c in sweep)
    for c in sweep:
        c.clear_samples()
    barrier()
    t_wall0 = time.perf_counter()
    evs = [one_step(sweep) for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = sum(c.launch_count for c in sweep) - launches0
    launches_per_rank = launches
    ms_steps = [a.elapsed_time(b) for a, b in evs]
    total_ms = sum(ms_steps)
    if dist is not None:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    attempts = float(n_sites) * PASSES_PER_STEP * args.steps * world
    value = attempts / (total_ms * 1e-3)
    # sanity: the run really sampled and moved
    i_tc = T_SWEEP.index(T_HEADLINE)
    S, B = sweep[i_tc].samples_sb(0)
    assert len(S) == PASSES_PER_STEP * args.steps
    x_mean = float((N0 * N1 + S.astype(np.float64)).mean() / 2.0 / (N0 * N1))
    acc_tc = sweep[i_tc].counters(0)
    main_variant = sweep[0].kernel_variant
    for c in sweep:
        c.close()

    # ---- the same eight lattices advanced concurrently by the HBM-streaming strip
    # kernel (one context, eight chains): the workload whose DRAM traffic is real
    lat.set_kernel_variant("bulk2d")
    for _ in range(2):
        one_step([lat])
    torch.cuda.synchronize()
    st_launch0 = lat.launch_count
    st_evs = [one_step([lat]) for _ in range(max(1, min(args.steps, 3)))]
    torch.cuda.synchronize()
    st_ms = sum(a.elapsed_time(b) for a, b in st_evs)
    st_launches = lat.launch_count - st_launch0
    st_value = float(n_sites) * PASSES_PER_STEP * len(st_evs) / (st_ms * 1e-3)
    st_bytes_per_launch = ALGO_BYTES_PER_ATTEMPT * float(n_sites) * PASSES_PER_STEP * len(st_evs) / max(1, st_launches)
    st_launch_s = st_ms * 1e-3 / max(1, st_launches)

    # ---- end to end through the C ABI with HOST buffers (H2D + D2H in the timed region)
    # The lattices of a temperature sweep are independent, so the end-to-end leg
    # drives them as E2E_GROUPS contexts (chains i*k .. i*k+k-1 each, same global
    # Philox streams through set_chain_offset) on their own CUDA streams: the
    # upload of one group overlaps the sweeps of another, and the downloads of the
    # first groups overlap the sweeps of the last.
    host_out = torch.empty((n_lat, per_lat), dtype=torch.int32).pin_memory()
    e2e_steps = max(1, min(args.steps, 5))
    n_groups = max(1, min(args.e2e_groups, n_lat))
    while n_lat % n_groups:
        n_groups -= 1
    per_group = n_lat // n_groups
    lat.close()
    groups = []
    for g in range(n_groups):
        st = torch.cuda.Stream()
        lg = IsingLatticeGPU([N0, N1], n_chains=per_group, device=local_rank, J=J)
        lg.set_stream(st.cuda_stream)
        for k in range(per_group):
            lg.set_conditions(T_SWEEP[g * per_group + k], mu, chain=k)
        lg.seed_philox(0xC0FFEE + rank)
        lg.set_chain_offset(g * per_group)
        if args.e2e_variant != "auto":
            lg.set_kernel_variant(args.e2e_variant)
        groups.append(lg)

    def enqueue(g):
        # inputs of one step for group g: H2D of the int32 occupation + colour-plane
        # split, then the sweeps -- all asynchronous on the group's stream
        lg = groups[g]
        for k in range(per_group):
            lg.upload(host_occ.numpy()[g * per_group + k], k)
        lg.clear_samples()
        lg.run_passes(PASSES_PER_STEP, MODE_CHECKERBOARD, 1)

    def collect(g):
        # results of one step for group g (blocks on the group's stream only)
        lg = groups[g]
        out = []
        for k in range(per_group):
            lg.download(k, out=host_out.numpy()[g * per_group + k])  # D2H of the final occupation
            out.append(lg.samples_sb(k))  # D2H of the sampled (S, B) series
        return out

    # Every step uploads all lattices, sweeps them and downloads all results.
    # The calls are ordered per group -- collect step s, enqueue step s+1 -- so
    # that while the host waits for one group's download the other groups'
    # copies and sweeps are already queued (H2D, D2H and compute overlap).
    for g in range(n_groups):
        enqueue(g)
    for g in range(n_groups):
        collect(g)
    barrier()
    t0 = time.perf_counter()
    for g in range(n_groups):
        enqueue(g)
    for s in range(e2e_steps):
        for g in range(n_groups):
            collect(g)
            if s + 1 < e2e_steps:
                enqueue(g)
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = float(n_sites) * PASSES_PER_STEP * e2e_steps * world / e2e_s
    e2e_launches = sum(lg.launch_count for lg in groups)
    for lg in groups:
        lg.close()

    # ---- the single 4096x4096 lattice at T_c on its own (tiled shared-memory kernel)
    single = IsingLatticeGPU([N0, N1], device=local_rank, J=J)
    single.set_stream(stream.cuda_stream)
    single.set_conditions(T_HEADLINE, mu)
    single.seed_philox(0xC0FFEE + rank)
    single.randomize(12345 + rank, 0.5)
    with torch.cuda.stream(stream):
        single.run_passes(60, MODE_CHECKERBOARD, 1)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        single.run_passes(PASSES_PER_STEP, MODE_CHECKERBOARD, 1)
        s1.record(stream)
    torch.cuda.synchronize()
    single_value = float(per_lat) * PASSES_PER_STEP / (s0.elapsed_time(s1) * 1e-3)
    single_variant = single.kernel_variant
    single.close()

    peak, peak_src = measured_peak()
    n_launches_step = launches_per_rank / max(1, args.steps)
    avg_launch_s = (sum(ms_steps) * 1e-3) / max(1, launches_per_rank)
    bytes_per_launch = ALGO_BYTES_PER_ATTEMPT * float(n_sites) * PASSES_PER_STEP * args.steps / max(1, launches_per_rank)
    achieved = bytes_per_launch / avg_launch_s / 1e9

    line = {
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u8 occupation, u32 Philox/threshold compare, f64 tables",
        "data": "synthetic",
        "config": workload_config(world),
        "clocks": clocks,
        "e2e": {
            "value": e2e_value,
            "unit": UNIT,
            "h2d_bytes_per_step": 4 * n_sites,
            "d2h_bytes_per_step": 4 * n_sites + 16 * PASSES_PER_STEP * n_lat,
            "steps": e2e_steps,
            "contexts": n_groups,
            "gpu_launches": e2e_launches,
            "note": "pinned int32 host buffers in and out through cmg_upload/download_occupation_i32; "
                    f"{n_groups} contexts of {per_group} lattice(s) on separate streams, calls ordered collect(step s) -> enqueue(step s+1) per context, "
                    "so H2D, D2H and sweeps of different contexts overlap; every step uploads and downloads every lattice",
        },
        "gpu_launches": launches,
        "roofline": {
            "bound": "hbm",
            "achieved": achieved,
            "peak": peak,
            "peak_source": peak_src,
            "unit": "GB/s",
            "frac": achieved / peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of one k_ring2d launch from the ncu
            # --set full capture in profiles/ncu_ring2d_r1h.txt: 16.8 MB read (the staging of the
            # 16 MiB lattice) + 7 KB written inside the launch window (the write-back stays dirty
            # in L2) -- the lattice lives in shared memory, so DRAM traffic is far BELOW the
            # algorithmic bytes; for the HBM-streaming form see streaming_8_lattices
            "traffic": 16.84e6 if main_variant == "ring2d" else (162.2e6 if main_variant == "bulk2d" else None),
            "kernel": ("k_halfsweep_" if main_variant.startswith("bulk") else "k_") + main_variant,
            "algorithmic_bytes_per_launch": bytes_per_launch,
            "avg_launch_us": avg_launch_s * 1e6,
            "launches_per_step": n_launches_step,
            "note": "3 B per attempted flip (int8, two colour planes) is the algorithmic HBM traffic of a half-sweep; k_ring2d keeps the lattice on chip for up to 128 passes per launch, so the fraction compares its rate with what an HBM-streaming kernel could at best do. The update is bound by integer issue (Philox4x32-10: FMA-heavy pipe 67 % busy, issue slots 60 %): see DESIGN.md / profiles/",
        },
        "streaming_8_lattices": {
            "value": st_value,
            "unit": UNIT,
            "kernel": "k_halfsweep_bulk2d",
            "workload": "the same eight lattices as eight chains of one context, advanced concurrently, every half-sweep streaming 201 MB through HBM (128 MiB of planes > L2)",
            "achieved": st_bytes_per_launch / st_launch_s / 1e9,
            "frac": st_bytes_per_launch / st_launch_s / 1e9 / peak,
            "avg_launch_us": st_launch_s * 1e6,
            "algorithmic_bytes_per_launch": st_bytes_per_launch,
            # profiles/ncu_bulk2d_sweep8_r1f.txt: 134.3 MB read + 27.9 MB written back within the
            # launch window (the rest of the 67 MB written stays dirty in L2 until the next launch)
            "traffic": 162.2e6,
        },
        "single_lattice": {"value": single_value, "unit": UNIT, "kernel": "k_" + single_variant, "workload": f"one {N0}x{N1} lattice at T=2633 K, sampling every pass", "frac_of_roofline": single_value * ALGO_BYTES_PER_ATTEMPT / 1e9 / peak},
        "wall_s_timed_region": t_wall,
        "check": {"T": T_HEADLINE, "mean_param_composition": x_mean, "acceptance_rate": acc_tc[1] / max(1, acc_tc[1] + acc_tc[2])},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        r = run_cpu_reference(N0, N1, T_HEADLINE, MU, 2, 1, False, cores)
        line["cpu_baseline"] = {
            "value": r["attempts_per_s"],
            "unit": UNIT,
            "cores": cores,
            "kind": "port",
            "sample": f"{cores} independent chains (one per core) x 2 passes of the {N0}x{N1} lattice, use_nlist=false, sampling every pass ({r['wall_s']:.1f} s wall)",
        }
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--value-variant", default="auto", help="kernel variant of the timed sweep (auto: k_ring2d)")
    ap.add_argument("--e2e-groups", type=int, default=8, help="contexts the end-to-end leg splits the lattices into")
    ap.add_argument("--e2e-variant", default="auto", help="kernel variant of the end-to-end contexts (auto: one lattice per context runs resident in shared memory, k_ring2d, while the other contexts copy)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
