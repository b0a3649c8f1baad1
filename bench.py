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
passes).  Reported next to it: the same eight lattices advanced concurrently by
the HBM-streaming strip kernel (`streaming_8_lattices`, k_halfsweep_bulk2d,
with its DRAM traffic) and a single lattice at T_c alone (`single_lattice`).

With N > 1 GPUs the workload is BASELINE.json configs[4]: ONE 65536x65536
lattice (4.3 G sites) at T_c cut into N column slabs, one per GPU, whose
half-sweep kernels push their two boundary columns straight into the
neighbours' halo buffers over NVLink (CUDA IPC peer stores + release/acquire
flags, no collective and no host work between half-sweeps; the pass loop is
cmg_slab_run_passes inside the C-ABI library).  A step is one config-5 run:
110 passes (10 + 100) sampling every 10.  Total work is fixed as N grows
("strong"); `halo_fraction` is the share of the step time the exchange costs
(same run with the exchange switched off), and an 8192^2 proxy is checked
bit-for-bit against a single-GPU run inside the same process group.  Extra
keys: the communication-free replica sweep of round 1 and the 1024-chain
256^2 (T, mu) grid (configs[3]) sharded over the ranks.

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
# N > 1: BASELINE.json configs[4]
SLAB_N = int(os.environ.get("CMG_BENCH_SLAB_N", "65536"))  # (override: small boxes / tests)
SLAB_PASSES_PER_STEP = 110  # 10 + 100 passes
SLAB_SAMPLE_PERIOD = 10
PROXY_N = 8192
GRID_SHAPE, GRID_POINTS = [256, 256], 32  # configs[3]: 32 x 32 (T, mu) grid of 256^2 chains


def ncu_traffic(kernel):
    """DRAM bytes (read + write) of one launch of `kernel`, parsed from the committed
    ncu --set full summary that profiles/traffic_sources.json names for it (mean
    over the launches the file holds): (bytes or None, "profiles/<file> (...)" or None)."""
    import re

    try:
        src = json.load(open(os.path.join(ROOT, "profiles", "traffic_sources.json")))[kernel]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        parts = {}
        for ln in open(os.path.join(ROOT, "profiles", src["file"])):
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                if ln.startswith(key + " ["):
                    unit = ln.split("[")[1].split("]")[0]
                    vals = [float(v) for v in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", ln.split("=", 1)[1])]
                    parts.setdefault(key, []).extend(v * scale[unit] for v in vals)
        if len(parts) < 2:
            return None, None
        total = sum(sum(v) / len(v) for v in parts.values())
        return total, "profiles/" + src["file"] + " (" + src.get("what", "") + ")"
    except Exception:
        return None, None


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
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    # one pass of the stated 4096x4096 lattice per core and step (~3-4 s): a bounded
    # sample of the workload.  At N > 1 the workload is the 65536^2 lattice, which the
    # reference cannot construct (its site count is an int product that wraps to 0,
    # include/casm/monte/ising_cpp/model.hh:28-29) and one chain per core could not
    # hold either: the arm then runs the same 4096^2 sample and says so (proxy: true).
    n0, n1, passes = N0, N1, 1
    sample = f"{cores} independent chains (one per core) x 1 pass of a {N0}x{N1} lattice per step, use_nlist=false, sampling every pass"
    cfg = workload_config(1) if world == 1 else slab_config(world)
    if world > 1:
        cfg = dict(cfg)
        cfg["proxy"] = True
        cfg["workload"] = f"PROXY for [{cfg['workload']}]: {sample} -- the reference path cannot hold a {SLAB_N}x{SLAB_N} lattice"
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
        "scaling": "weak" if world == 1 else "strong",
        "vs_baseline": None,
        "dtype": "int32 occupation / f64 energies (CPU)",
        "data": "synthetic",
        "config": cfg,
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
        "proxy": False,
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


def slab_config(n_gpus):
    return {
        "workload": f"2D Ising SGC {SLAB_N}x{SLAB_N} ({SLAB_N * SLAB_N / 1e9:.1f} G sites) at T_c=2633 K, mu=0, J=0.1 eV, domain-decomposed into "
        f"{n_gpus} column slabs over {n_gpus} B200 with per-half-sweep NVLink halo exchange fused into the half-sweep kernel "
        "(peer stores of the two boundary columns + release/acquire flags; pass loop inside the C-ABI library, cmg_slab_run_passes)",
        "lattice": [SLAB_N, SLAB_N],
        "slabs": n_gpus,
        "passes_per_step": SLAB_PASSES_PER_STEP,
        "sample_period": SLAB_SAMPLE_PERIOD,
        "initial_state": "i.i.d. +1/-1 (Philox keyed on global site indices, seed 12345)",
        "philox_seed": "0xC0FFEE (counters keyed on global site indices: the trajectory does not depend on the number of GPUs)",
        "l2": f"inputs larger than L2: every half-sweep streams its slab ({SLAB_N * SLAB_N // n_gpus / 2**20:.0f} MiB of int8 planes per GPU) from HBM",
        "timing": "CUDA events on the launching stream per step, summed; max over ranks",
        "proxy": False,
    }


def ours(args):
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return ours_slab(args)
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
    host_occ = torch.empty((n_lat, per_lat), dtype=torch.int8).pin_memory()
    sweep = []
    for k, T in enumerate(T_SWEEP):
        lat.download_i8(k, out=host_occ.numpy()[k])
        lk = IsingLatticeGPU([N0, N1], device=local_rank, J=J)
        lk.set_stream(stream.cuda_stream)
        lk.set_conditions(T, mu)
        lk.seed_philox(0xC0FFEE + rank)
        lk.set_chain_offset(k)
        if args.value_variant != "auto":
            lk.set_kernel_variant(args.value_variant)
        lk.upload_i8(host_occ.numpy()[k])
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
    host_out = torch.empty((n_lat, per_lat), dtype=torch.int8).pin_memory()
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
        # inputs of one step for group g: H2D of the int8 occupation + colour-plane
        # split, then the sweeps -- all asynchronous on the group's stream
        lg = groups[g]
        for k in range(per_group):
            lg.upload_i8(host_occ.numpy()[g * per_group + k], k)
        lg.clear_samples()
        lg.run_passes(PASSES_PER_STEP, MODE_CHECKERBOARD, 1)

    def collect(g):
        # results of one step for group g (blocks on the group's stream only)
        lg = groups[g]
        out = []
        for k in range(per_group):
            lg.download_i8(k, out=host_out.numpy()[g * per_group + k])  # D2H of the final occupation
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
    # the same lattice with the opt-in seven-round Philox stream (cmg_set_philox_rounds(7): the
    # fewest rounds that pass BigCrush; NOT what `value` is measured with)
    single.set_philox_rounds(7)
    with torch.cuda.stream(stream):
        single.run_passes(60, MODE_CHECKERBOARD, 1)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        single.run_passes(PASSES_PER_STEP, MODE_CHECKERBOARD, 1)
        s1.record(stream)
    torch.cuda.synchronize()
    philox7_value = float(per_lat) * PASSES_PER_STEP / (s0.elapsed_time(s1) * 1e-3)
    single.close()

    # ---- the N > 1 workload (configs[4], the 65536^2 lattice) on this one GPU, undecomposed:
    # the single-GPU base of the strong-scaling series bench.py --gpus N reports
    big = IsingLatticeGPU([SLAB_N, SLAB_N], device=local_rank, J=J)
    big.set_stream(stream.cuda_stream)
    big.set_conditions(T_HEADLINE, MU)
    big.seed_philox(0xC0FFEE)
    big.randomize(12345, 0.5)
    with torch.cuda.stream(stream):
        big.run_passes(2, MODE_CHECKERBOARD, 0)
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record(stream)
        big.run_passes(20, MODE_CHECKERBOARD, SLAB_SAMPLE_PERIOD)
        b1.record(stream)
    torch.cuda.synchronize()
    big_value = float(SLAB_N) * SLAB_N * 20 / (b0.elapsed_time(b1) * 1e-3)
    big_variant = big.kernel_variant
    big.close()

    peak, peak_src = measured_peak()
    main_kernel = ("k_halfsweep_" if main_variant.startswith("bulk") else "k_") + main_variant
    main_traffic, main_traffic_src = ncu_traffic(main_kernel)
    st_traffic, st_traffic_src = ncu_traffic("k_halfsweep_bulk2d")
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
            "h2d_bytes_per_step": n_sites,
            "d2h_bytes_per_step": n_sites + 16 * PASSES_PER_STEP * n_lat,
            "steps": e2e_steps,
            "contexts": n_groups,
            "gpu_launches": e2e_launches,
            "note": "pinned int8 (+1/-1) host buffers in and out through cmg_upload/download_occupation_i8 (the int32 form of the reference's Eigen::VectorXi is cmg_upload/download_occupation_i32, 4x the PCIe bytes); "
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
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel, parsed
            # from the committed ncu --set full summary named in traffic_source.  k_ring2d keeps
            # the lattice in shared memory, so its DRAM traffic (the staging of the 16 MiB
            # lattice) is far BELOW the algorithmic bytes; the HBM-streaming form is
            # streaming_8_lattices
            "traffic": main_traffic,
            "traffic_source": main_traffic_src,
            "kernel": main_kernel,
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
            # read + written back within the launch window (the rest of the 67 MB written stays
            # dirty in L2 until the next launch)
            "traffic": st_traffic,
            "traffic_source": st_traffic_src,
        },
        "single_lattice_philox7_opt_in": {"value": philox7_value, "unit": UNIT, "kernel": "k_" + single_variant, "workload": f"one {N0}x{N1} lattice at T=2633 K, sampling every pass, Philox4x32-7 instead of the default Philox4x32-10 (opt-in, cmg_set_philox_rounds)", "frac_of_roofline": philox7_value * ALGO_BYTES_PER_ATTEMPT / 1e9 / peak},
        "single_lattice": {"value": single_value, "unit": UNIT, "kernel": "k_" + single_variant, "workload": f"one {N0}x{N1} lattice at T=2633 K, sampling every pass", "frac_of_roofline": single_value * ALGO_BYTES_PER_ATTEMPT / 1e9 / peak},
        "decomposed_lattice_on_one_gpu": {"value": big_value, "unit": UNIT, "kernel": "k_halfsweep_" + big_variant, "workload": f"the {SLAB_N}x{SLAB_N} lattice of the N > 1 runs (configs[4]) undecomposed on this GPU, 20 passes sampling every {SLAB_SAMPLE_PERIOD}, streaming {SLAB_N * SLAB_N / 2**30:.0f} GiB of planes from HBM", "frac_of_roofline": big_value * ALGO_BYTES_PER_ATTEMPT / 1e9 / peak},
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


def ours_slab(args):
    """N > 1: the 65536^2 lattice decomposed into column slabs (configs[4])."""
    import hashlib

    import numpy as np
    import torch
    import torch.distributed as dist

    from casmcode_monte_b200 import MODE_CHECKERBOARD, IsingLatticeGPU
    from casmcode_monte_b200.parallel import GpuSlabEngine, SlabRing, shard_chains, slab_columns

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    seed = 0xC0FFEE

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], device="cuda", dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    def make_ring(n, T, mu, randomize_seed=None, occ=None, n1=None, variant=None):
        n1 = n if n1 is None else n1
        cb, nc = slab_columns(n1, world, rank)
        eng = GpuSlabEngine([n, n1], cb, nc, J, T, mu, seed, device=local_rank, stream=stream.cuda_stream)
        if variant:
            eng.lat.set_kernel_variant(variant)
        if occ is not None:
            eng.upload(occ[n * cb : n * (cb + nc)])
        else:
            eng.lat.randomize(randomize_seed, 0.5)  # keyed on global site indices
        eng.sync()
        with torch.cuda.stream(stream):
            ring = SlabRing(eng, rank, world, dist, transport="peer")
            ring.prime()  # halos from the neighbours (NCCL once), then CUDA-IPC attach
        barrier()
        return eng, ring, (cb, nc)

    # ---- bit-identity of an 8192^2 proxy against ONE GPU, inside this process group
    proxy_passes = 6
    rng = np.random.default_rng(20261018)
    proxy_full = rng.choice(np.array([-1, 1], dtype=np.int32), size=PROXY_N * PROXY_N)
    eng, ring, (cb, nc) = make_ring(PROXY_N, T_HEADLINE, 0.01, occ=proxy_full)
    with torch.cuda.stream(stream):
        ring.run_passes(proxy_passes, sample_period=2)
    eng.sync()
    digest = hashlib.sha256(np.ascontiguousarray(eng.lat.download_i8()).tobytes()).hexdigest()
    S_part, B_part = eng.lat.samples_sb(0)
    parts = [None] * world
    dist.all_gather_object(parts, (digest, S_part.tolist(), B_part.tolist(), cb, nc))
    eng.lat.close()
    proxy = None
    if rank == 0:
        ref = IsingLatticeGPU([PROXY_N, PROXY_N], device=local_rank, J=J)
        ref.set_conditions(T_HEADLINE, 0.01)
        ref.seed_philox(seed)
        ref.upload(proxy_full)
        ref.run_passes(proxy_passes, MODE_CHECKERBOARD, 2)
        full = ref.download_i8()
        Sr, Br = ref.samples_sb(0)
        same = all(
            hashlib.sha256(np.ascontiguousarray(full[PROXY_N * c0 : PROXY_N * (c0 + n_)]).tobytes()).hexdigest() == d
            for d, _, _, c0, n_ in parts
        )
        # a slab samples `ones` and B of its own columns; S = 2*ones - N is affine, so the
        # slab values (each 2*ones_r - N_r) add up to the lattice's
        S_sum = np.sum([p[1] for p in parts], axis=0)
        B_sum = np.sum([p[2] for p in parts], axis=0)
        proxy = {
            "lattice": [PROXY_N, PROXY_N],
            "passes": proxy_passes,
            "single_gpu_kernel": ("k_halfsweep_" if ref.kernel_variant.startswith("bulk") else "k_") + ref.kernel_variant,
            "occupation_bit_identical_to_single_gpu": bool(same),
            "sampled_S_B_identical": bool(np.array_equal(S_sum, Sr) and np.array_equal(B_sum, Br)),
        }
        ref.close()
        assert proxy["occupation_bit_identical_to_single_gpu"] and proxy["sampled_S_B_identical"], proxy
    # ---- the same proxy RESIDENT in shared memory over the GPUs (k_ring2d on slabs: the outer
    # tiles of neighbouring GPUs trade their edge columns through peer memory inside the
    # cooperative kernel, one launch per GPU per 128 passes): bit-identity against the
    # streaming run above, then its rate.  8192 x 8192 needs four GPUs' shared memory; two
    # GPUs take 8192 x 4096.
    ring_n1 = PROXY_N if world >= 4 else PROXY_N // 2
    ring_occ = proxy_full[: PROXY_N * ring_n1]
    ring_err = []

    def guarded(f):
        # every rank walks the same sequence of collectives whatever happens to its own calls
        if not ring_err:
            try:
                return f()
            except Exception as exc:
                ring_err.append(str(exc)[:300])
        return None

    def timed_passes(ring, n_passes, reps):
        evs = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
                guarded(lambda: ring.run_passes(n_passes, sample_period=10))
                e1.record(stream)
            evs.append((e0, e1))
        barrier()
        return max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))

    def digest_after(ring, eng):
        with torch.cuda.stream(stream):
            guarded(lambda: ring.run_passes(proxy_passes, sample_period=2))
        guarded(eng.sync)
        d = guarded(lambda: hashlib.sha256(np.ascontiguousarray(eng.lat.download_i8()).tobytes()).hexdigest())
        barrier()
        return d

    ring_passes, ring_reps = 384, 3
    eng, ring, _ = make_ring(PROXY_N, T_HEADLINE, 0.01, occ=ring_occ, n1=ring_n1, variant="ring2d")
    r_digest = digest_after(ring, eng)
    r_kernel = eng.lat.kernel_variant
    timed_passes(ring, ring_passes, 1)  # warm-up
    r_ms = timed_passes(ring, ring_passes, ring_reps)
    guarded(eng.sync)
    eng.lat.close()
    # the streaming run of the same lattice on the same ranks
    eng, ring, _ = make_ring(PROXY_N, T_HEADLINE, 0.01, occ=ring_occ, n1=ring_n1)
    s_digest = digest_after(ring, eng)
    s_ms = timed_passes(ring, 40, 2)
    eng.lat.close()
    same = sum_over_ranks(0 if (r_digest is not None and r_digest == s_digest) else 1) == 0
    n_failed = sum_over_ranks(1 if ring_err else 0)
    resident = {
        "lattice": [PROXY_N, ring_n1],
        "value": float(PROXY_N) * ring_n1 * ring_passes * ring_reps / (r_ms * 1e-3) if n_failed == 0 else None,
        "unit": "attempts/s",
        "kernel": "k_" + r_kernel,
        "passes_per_launch": 128,
        "streaming_same_lattice": float(PROXY_N) * ring_n1 * 40 * 2 / (s_ms * 1e-3),
        "bit_identical_to_the_streaming_run": bool(same),
        "ranks_with_errors": n_failed,
        "error": ring_err[0] if ring_err else None,
        "workload": "the lattice resident in the GPUs' shared memory (one cooperative k_ring2d launch per GPU per 128 passes); outer tiles of neighbouring GPUs exchange edge columns by stores into each other's mailboxes over NVLink inside the kernel; sampling every 10 passes",
    }
    del proxy_full
    barrier()

    # ---- the timed run: one 65536^2 lattice over `world` slabs
    n = SLAB_N
    eng, ring, (cb, nc) = make_ring(n, T_HEADLINE, MU, randomize_seed=12345)
    lat = eng.lat
    slab_sites = n * nc

    def one_step():
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            ring.run_passes(SLAB_PASSES_PER_STEP, SLAB_SAMPLE_PERIOD)
            e1.record(stream)
        return (e0, e1)

    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    lat.clear_samples()
    launches0 = lat.launch_count
    barrier()
    t_wall0 = time.perf_counter()
    evs = [one_step() for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches_rank = lat.launch_count - launches0
    ms_rank = sum(a.elapsed_time(b) for a, b in evs)
    total_ms = max_over_ranks(ms_rank)
    launches = sum_over_ranks(launches_rank)
    attempts = float(n) * n * SLAB_PASSES_PER_STEP * args.steps
    value = attempts / (total_ms * 1e-3)
    S_part, B_part = lat.samples_sb(0)
    assert len(S_part) == (SLAB_PASSES_PER_STEP // SLAB_SAMPLE_PERIOD) * args.steps
    acc = lat.counters(0)
    acc_tot, att_tot = sum_over_ranks(int(acc[1])), sum_over_ranks(int(acc[1] + acc[2]))
    S_last = sum_over_ranks(int(S_part[-1]))

    # ---- end to end: host buffers in and out every step (bit-packed occupation of the slab
    # from pinned memory, halos refreshed from the neighbours, the 110 passes, bit-packed
    # result + the sampled series back), wall clock between barriers, max over ranks
    host_bits = torch.empty((slab_sites + 7) // 8, dtype=torch.uint8).pin_memory()
    host_out = torch.empty_like(host_bits).pin_memory()
    lat.download_bits(out=host_bits.numpy())
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_step():
        lat.upload_bits(host_bits.numpy())
        lat.clear_samples()
        with torch.cuda.stream(stream):
            for colour in (0, 1):
                ring.exchange(colour)  # the new state's boundary columns -> neighbours' halos
            ring.run_passes(SLAB_PASSES_PER_STEP, SLAB_SAMPLE_PERIOD)
        lat.download_bits(out=host_out.numpy())
        return lat.samples_sb(0)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = float(n) * n * SLAB_PASSES_PER_STEP * e2e_steps / e2e_s

    # ---- share of the step the halo exchange costs: the same run with the exchange off
    # (kernels neither push their boundary columns nor wait for the neighbours; the
    # trajectory is meaningless from here on, only the time is used)
    barrier()
    lat.slab_set_halo_exchange(False)
    off_steps = max(1, min(args.steps, 3))
    one_step()
    barrier()
    off_evs = [one_step() for _ in range(off_steps)]
    barrier()
    off_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in off_evs) / off_steps)
    on_ms = total_ms / args.steps
    halo_fraction = max(0.0, 1.0 - off_ms / on_ms)
    lat.close()
    barrier()

    # ---- extra: the communication-free replica sweep of round 1 (each rank the eight
    # 4096^2 temperature lattices at its own mu), one step
    mu_r = MU_GRID[rank % len(MU_GRID)]
    rep = []
    for k, T in enumerate(T_SWEEP):
        lk = IsingLatticeGPU([N0, N1], device=local_rank, J=J)
        lk.set_stream(stream.cuda_stream)
        lk.set_conditions(T, mu_r)
        lk.seed_philox(seed + rank)
        lk.set_chain_offset(k)
        lk.randomize(12345 + rank + 97 * k, 0.5)
        rep.append(lk)

    def replica_step():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for c in rep:
                c.run_passes(PASSES_PER_STEP, MODE_CHECKERBOARD, 1)
            e1.record(stream)
        return e0, e1

    replica_step()
    barrier()
    r_evs = [replica_step() for _ in range(2)]
    barrier()
    rep_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in r_evs))
    rep_value = float(N0) * N1 * len(T_SWEEP) * PASSES_PER_STEP * 2 * world / (rep_ms * 1e-3)
    rep_kernel = rep[0].kernel_variant
    for c in rep:
        c.close()

    # ---- extra: configs[3], the 1024-point (T, mu) grid of 256^2 chains sharded over the ranks
    Ts = np.linspace(1500.0, 4000.0, GRID_POINTS)
    mus = np.linspace(-0.2, 0.2, GRID_POINTS)
    conditions = [(float(T), float(m)) for T in Ts for m in mus]
    mine = shard_chains(len(conditions), world, rank)
    g = IsingLatticeGPU(GRID_SHAPE, n_chains=len(mine), device=local_rank, J=J)
    g.set_stream(stream.cuda_stream)
    for loc, gi in enumerate(mine):
        g.set_conditions(conditions[gi][0], conditions[gi][1], chain=loc)
    g.seed_philox(seed)
    g.set_chain_offset(mine[0])
    grid_equil, grid_passes = 2000, 8000
    with torch.cuda.stream(stream):
        g.run_passes(200, MODE_CHECKERBOARD, 0)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        g.run_passes(grid_equil - 200, MODE_CHECKERBOARD, 0)
        g.run_passes(grid_passes, MODE_CHECKERBOARD, 1)
        e1.record(stream)
    barrier()
    grid_ms = max_over_ranks(e0.elapsed_time(e1))
    grid_value = float(GRID_SHAPE[0]) * GRID_SHAPE[1] * len(conditions) * (grid_equil - 200 + grid_passes) / (grid_ms * 1e-3)
    grid_kernel = g.kernel_variant
    grid_samples = g.n_samples
    g.close()

    peak, peak_src = measured_peak()
    bytes_per_launch = ALGO_BYTES_PER_ATTEMPT * float(slab_sites) / 2.0  # one half-sweep of one slab
    avg_launch_s = (ms_rank * 1e-3) / max(1, launches_rank)
    achieved = bytes_per_launch / avg_launch_s / 1e9
    traffic, traffic_src = ncu_traffic(f"k_halfsweep_bulk2d@slab{world}")
    line = {
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "u8 occupation, u32 Philox/threshold compare, f64 tables",
        "data": "synthetic",
        "config": slab_config(world),
        "clocks": clocks,
        "halo_fraction": halo_fraction,
        "halo": {
            "ms_per_step_with_exchange": on_ms,
            "ms_per_step_exchange_off": off_ms,
            "bytes_per_half_sweep_per_gpu": 2 * (n // 2),
            "transport": "peer stores into the neighbours' halo buffers (CUDA IPC over NVLink) fused into k_halfsweep_bulk2d, release/acquire flags, bounded waits",
            "how": "halo_fraction = 1 - (step time with cmg_slab_set_halo_exchange(0)) / (step time with the exchange), same lattice, max over ranks",
        },
        "proxy_bit_identity": proxy,
        "resident_ring_8192": resident,
        "e2e": {
            "value": e2e_value,
            "unit": UNIT,
            "h2d_bytes_per_step": ((slab_sites + 7) // 8) * world,
            "d2h_bytes_per_step": ((slab_sites + 7) // 8 + 16 * (SLAB_PASSES_PER_STEP // SLAB_SAMPLE_PERIOD)) * world,
            "steps": e2e_steps,
            "note": "every step: each rank uploads its slab from pinned host memory as one bit per site (cmg_upload_occupation_bits), "
                    "the boundary columns of the new state go to the neighbours' halos (NCCL send/recv), the 110 passes run, and the "
                    "bit-packed slab plus the sampled (S, B) series come back (cmg_download_occupation_bits, cmg_read_samples_sb); wall clock between barriers, max over ranks",
        },
        "gpu_launches": launches,
        "roofline": {
            "bound": "hbm",
            "achieved": achieved,
            "peak": peak,
            "peak_source": peak_src,
            "unit": "GB/s",
            "frac": achieved / peak,
            "traffic": traffic,
            "traffic_source": traffic_src,
            "kernel": "k_halfsweep_bulk2d",
            "algorithmic_bytes_per_launch": bytes_per_launch,
            "avg_launch_us": avg_launch_s * 1e6,
            "launches_per_step": launches_rank / max(1, args.steps),
            "note": "per GPU: one launch = one coloured half-sweep of this rank's slab, 3 B per attempted flip; avg launch time = this rank's CUDA-event time / its launches (rank 0)",
        },
        "replica_sweep": {
            "value": rep_value,
            "unit": UNIT,
            "kernel": "k_" + rep_kernel,
            "workload": f"round 1's N>1 workload: every rank the eight {N0}x{N1} temperature lattices at its own mu_r, no communication, 2 steps of {PASSES_PER_STEP} passes",
        },
        "chain_grid_1024x256": {
            "value": grid_value,
            "unit": UNIT,
            "kernel": "k_" + grid_kernel,
            "workload": f"BASELINE configs[3]: {len(conditions)} independent 256x256 chains (32 x 32 grid, T in [1500, 4000] K, mu in [-0.2, 0.2] eV) sharded over {world} GPUs "
                        f"({len(mine)} chains on rank 0), {grid_equil} + {grid_passes} passes, sampling every pass after equilibration, no communication",
            "samples_per_chain": grid_samples,
        },
        "wall_s_timed_region": t_wall,
        "check": {"T": T_HEADLINE, "acceptance_rate": acc_tot / max(1, att_tot), "mean_param_composition_last_sample": (float(n) * n + S_last) / 2.0 / (float(n) * n)},
    }
    if rank == 0:
        print(json.dumps(line), flush=True)
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
