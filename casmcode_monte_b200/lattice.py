"""IsingLatticeGPU: numpy-facing wrapper of one C-ABI context.

Thin by design: every method is one or two calls of include/casm_monte_gpu.h.
The reference-shaped classes (IsingConfiguration, SemiGrandCanonicalCalculator,
...) are built on top of this in casmcode_monte_b200.monte.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import (  # noqa: F401
    KB,
    MODE_CHECKERBOARD,
    MODE_SERIAL_REFERENCE,
    Q_FORMATION_ENERGY,
    Q_PARAM_COMPOSITION,
    Q_POTENTIAL_ENERGY,
    CmgError,
    check,
)


def _p(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


class IsingLatticeGPU:
    """n_chains lattices of one shape on one GPU (cmg_create)."""

    def __init__(self, shape, n_chains=1, device=0, J=None, slab=None):
        self._lib = _capi.load()
        self.shape = tuple(int(s) for s in shape)
        self.dim = len(self.shape)
        self.n_chains = int(n_chains)
        self.device = int(device)
        self._ctx = C.c_void_p()
        sh = (C.c_int64 * max(3, self.dim))(*(list(self.shape) + [1] * (3 - self.dim)))
        if slab is None:
            check(self._lib.cmg_create(self.dim, sh, self.n_chains, self.device, C.byref(self._ctx)))
            self.local_shape = self.shape
        else:
            col_begin, n_cols = slab
            check(self._lib.cmg_create_slab(self.dim, sh, int(col_begin), int(n_cols), self.device, C.byref(self._ctx)))
            self.local_shape = (self.shape[0], int(n_cols))
        n = C.c_int64()
        check(self._lib.cmg_n_sites(self._ctx, C.byref(n)), self._ctx)
        self.n_sites = n.value
        if J is not None:
            self.set_model(J)

    # -- lifecycle --
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._lib.cmg_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        check(rc, self._ctx)

    def set_stream(self, cuda_stream):
        self._ck(self._lib.cmg_set_stream(self._ctx, C.c_void_p(int(cuda_stream))))

    def sync(self):
        self._ck(self._lib.cmg_sync(self._ctx))

    # -- model / conditions --
    def set_model(self, J, lattice_type=1):
        self._ck(self._lib.cmg_set_model(self._ctx, float(J), int(lattice_type)))

    def set_conditions(self, temperature, mu, chain=-1):
        self._ck(self._lib.cmg_set_conditions(self._ctx, int(chain), float(temperature), float(mu)))

    def tables(self, chain=0):
        n = 2 * (2 * self.dim + 1)
        dE = np.zeros(n)
        prob = np.zeros(n)
        thr = np.zeros(n, dtype=np.uint32)
        self._ck(self._lib.cmg_get_tables(self._ctx, chain, _p(dE, C.c_double), _p(prob, C.c_double), _p(thr, C.c_uint32)))
        return dE, prob, thr

    # -- occupation --
    def upload(self, occ, chain=0):
        occ = np.ascontiguousarray(occ, dtype=np.int32).ravel()
        self._ck(self._lib.cmg_upload_occupation_i32(self._ctx, chain, _p(occ, C.c_int32), occ.size))

    def download(self, chain=0, out=None):
        if out is None:
            out = np.empty(self.n_sites, dtype=np.int32)
        self._ck(self._lib.cmg_download_occupation_i32(self._ctx, chain, _p(out, C.c_int32), out.size))
        return out

    def upload_i8(self, occ, chain=0):
        """int8 +1/-1 per site (a quarter of the int32 form's PCIe traffic)."""
        occ = np.ascontiguousarray(occ, dtype=np.int8).ravel()
        self._ck(self._lib.cmg_upload_occupation_i8(self._ctx, chain, _p(occ, C.c_int8), occ.size))

    def download_i8(self, chain=0, out=None):
        if out is None:
            out = np.empty(self.n_sites, dtype=np.int8)
        self._ck(self._lib.cmg_download_occupation_i8(self._ctx, chain, _p(out, C.c_int8), out.size))
        return out

    def upload_bits(self, bits, chain=0):
        """One bit per site, as numpy.packbits(occ > 0, bitorder='little')."""
        bits = np.ascontiguousarray(bits, dtype=np.uint8).ravel()
        if bits.size != (self.n_sites + 7) // 8:
            raise ValueError("Error in set_occupation: size mismatch")
        self._ck(self._lib.cmg_upload_occupation_bits(self._ctx, chain, _p(bits, C.c_uint8), self.n_sites))

    def download_bits(self, chain=0, out=None):
        if out is None:
            out = np.empty((self.n_sites + 7) // 8, dtype=np.uint8)
        self._ck(self._lib.cmg_download_occupation_bits(self._ctx, chain, _p(out, C.c_uint8), self.n_sites))
        return out

    def upload_dev(self, dev_ptr, n, chain=0):
        self._ck(self._lib.cmg_upload_occupation_i32_dev(self._ctx, chain, C.c_void_p(int(dev_ptr)), int(n)))

    def download_dev(self, dev_ptr, n, chain=0):
        self._ck(self._lib.cmg_download_occupation_i32_dev(self._ctx, chain, C.c_void_p(int(dev_ptr)), int(n)))

    def get_occ(self, l, chain=0):
        v = C.c_int32()
        self._ck(self._lib.cmg_get_occ(self._ctx, chain, int(l), C.byref(v)))
        return v.value

    def set_occ(self, l, value, chain=0):
        self._ck(self._lib.cmg_set_occ(self._ctx, chain, int(l), int(value)))

    def event_delta(self, linear_site_index, new_occ, chain=0):
        """(dE_formation, dNx) of an event, model.hh:354-379 / :425-435."""
        ls = np.ascontiguousarray(linear_site_index, dtype=np.int64)
        no = np.ascontiguousarray(new_occ, dtype=np.int32)
        dE, dN = C.c_double(), C.c_double()
        self._ck(self._lib.cmg_event_delta(self._ctx, chain, ls.size, _p(ls, C.c_int64), _p(no, C.c_int32), C.byref(dE), C.byref(dN)))
        return dE.value, dN.value

    def fill(self, value, chain=-1):
        self._ck(self._lib.cmg_fill_occupation(self._ctx, chain, int(value)))

    def randomize(self, seed, p_up=0.5, chain=-1):
        self._ck(self._lib.cmg_randomize_occupation(self._ctx, chain, int(seed), float(p_up)))

    # -- rng --
    def seed_philox(self, seed):
        self._ck(self._lib.cmg_seed_philox(self._ctx, int(seed)))

    def set_philox_rounds(self, rounds):
        """10 (default) or 7 (the fewest rounds of Philox4x32 that pass BigCrush; opt-in)."""
        self._ck(self._lib.cmg_set_philox_rounds(self._ctx, int(rounds)))

    def set_pass_counter(self, t):
        self._ck(self._lib.cmg_set_pass_counter(self._ctx, int(t)))

    def set_chain_offset(self, global_index_of_chain_0):
        self._ck(self._lib.cmg_set_chain_offset(self._ctx, int(global_index_of_chain_0)))

    def seed_mt19937_64(self, seed, chain=0):
        self._ck(self._lib.cmg_seed_mt19937_64(self._ctx, chain, int(seed)))

    def set_engine_state(self, state312, position, chain=0):
        st = np.ascontiguousarray(state312, dtype=np.uint64)
        assert st.size == 312
        self._ck(self._lib.cmg_set_mt19937_64_state(self._ctx, chain, _p(st, C.c_uint64), int(position)))

    def get_engine_state(self, chain=0):
        st = np.zeros(312, dtype=np.uint64)
        pos = C.c_int()
        self._ck(self._lib.cmg_get_mt19937_64_state(self._ctx, chain, _p(st, C.c_uint64), C.byref(pos)))
        return st, pos.value

    def load_engine_dump(self, text, chain=0):
        """Accepts what operator<< of std::mt19937_64 prints (RandomNumberEngine.dump())."""
        w = [int(t) for t in text.split()]
        assert len(w) == 313
        self.set_engine_state(np.array(w[:312], dtype=np.uint64), w[312], chain)

    def engine_dump(self, chain=0):
        st, pos = self.get_engine_state(chain)
        return " ".join(str(int(v)) for v in st) + " " + str(pos)

    def rng_draw(self, requests, chain=0):
        """requests: list of ('int', max) / ('real', max); returns list of draws."""
        n = len(requests)
        imax = np.frombuffer((C.c_uint64 * n)(*[int(m) if k == "int" else 0 for k, m in requests]), dtype=np.uint64).view(np.int64).copy()
        rmax = np.array([float(m) if k == "real" else 0.0 for k, m in requests], dtype=np.float64)
        isr = np.array([1 if k == "real" else 0 for k, _ in requests], dtype=np.uint8)
        iout = np.zeros(n, dtype=np.int64)
        rout = np.zeros(n, dtype=np.float64)
        self._ck(
            self._lib.cmg_rng_draw(
                self._ctx, chain, n, _p(imax, C.c_int64), _p(rmax, C.c_double), _p(isr, C.c_uint8), _p(iout, C.c_int64), _p(rout, C.c_double)
            )
        )
        return [float(rout[i]) if isr[i] else int(iout.view(np.uint64)[i]) for i in range(n)]

    # -- stepping --
    def run_passes(self, n_passes, mode=MODE_CHECKERBOARD, sample_period=0):
        self._ck(self._lib.cmg_run_passes(self._ctx, int(n_passes), int(mode), int(sample_period)))

    def counters(self, chain=0):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self._ck(self._lib.cmg_counters(self._ctx, chain, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def reset_counters(self):
        self._ck(self._lib.cmg_reset_counters(self._ctx))

    # -- sampling --
    def sample_now(self, chain=0):
        S, B = C.c_int64(), C.c_int64()
        self._ck(self._lib.cmg_sample_now(self._ctx, chain, C.byref(S), C.byref(B)))
        return S.value, B.value

    def line_dots(self, chain=0):
        r = np.zeros(self.shape[0], dtype=np.int64)
        c = np.zeros(self.shape[1], dtype=np.int64)
        self._ck(self._lib.cmg_line_dots(self._ctx, chain, _p(r, C.c_int64), _p(c, C.c_int64)))
        return r, c

    @property
    def n_samples(self):
        n = C.c_int64()
        self._ck(self._lib.cmg_n_samples(self._ctx, C.byref(n)))
        return n.value

    def clear_samples(self):
        self._ck(self._lib.cmg_clear_samples(self._ctx))

    def samples_sb(self, chain=0, first=0, count=None):
        if count is None:
            count = self.n_samples - first
        S = np.zeros(count, dtype=np.int64)
        B = np.zeros(count, dtype=np.int64)
        self._ck(self._lib.cmg_read_samples_sb(self._ctx, chain, first, count, _p(S, C.c_int64), _p(B, C.c_int64)))
        return S, B

    def samples(self, quantity, chain=0, first=0, count=None):
        if count is None:
            count = self.n_samples - first
        out = np.zeros(count, dtype=np.float64)
        self._ck(self._lib.cmg_read_samples(self._ctx, chain, quantity, first, count, _p(out, C.c_double)))
        return out

    # -- probes --
    def delta_e_probe(self, chain=0):
        out = np.zeros(self.n_sites, dtype=np.float64)
        self._ck(self._lib.cmg_delta_e_probe(self._ctx, chain, _p(out, C.c_double)))
        return out

    def accept_probe(self, uniforms, chain=0):
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        assert u.size == self.n_sites
        out = np.zeros(self.n_sites, dtype=np.uint8)
        self._ck(self._lib.cmg_accept_probe(self._ctx, chain, _p(u, C.c_double), _p(out, C.c_uint8)))
        return out

    # -- statistics --
    def series_stats(self, quantity, chain=0, first=0, count=None, confidence=0.95):
        if count is None:
            count = self.n_samples - first
        m, p, v, k = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        self._ck(self._lib.cmg_series_stats(self._ctx, chain, quantity, first, count, confidence, C.byref(m), C.byref(p), C.byref(v), C.byref(k)))
        return {"mean": m.value, "calculated_precision": p.value, "variance": v.value, "k_star": k.value}

    def series_equilibration(self, quantity, abs_precision, chain=0, count=None):
        if count is None:
            count = self.n_samples
        e, n = C.c_int(), C.c_int64()
        self._ck(self._lib.cmg_series_equilibration(self._ctx, chain, quantity, count, abs_precision, C.byref(e), C.byref(n)))
        return bool(e.value), n.value

    def series_stats_all(self, quantity, first=None, count_total=None, confidence=0.95):
        if count_total is None:
            count_total = self.n_samples
        nc = self.n_chains
        m, p, v = np.zeros(nc), np.zeros(nc), np.zeros(nc)
        k = np.zeros(nc, dtype=np.int64)
        fp = None
        if first is not None:
            first = np.ascontiguousarray(first, dtype=np.int64)
            fp = _p(first, C.c_int64)
        self._ck(self._lib.cmg_series_stats_all(self._ctx, quantity, fp, count_total, confidence, _p(m, C.c_double), _p(p, C.c_double), _p(v, C.c_double), _p(k, C.c_int64)))
        return m, p, v, k

    def series_equilibration_all(self, quantity, abs_precision, count=None):
        if count is None:
            count = self.n_samples
        nc = self.n_chains
        e = np.zeros(nc, dtype=np.int32)
        n = np.zeros(nc, dtype=np.int64)
        self._ck(self._lib.cmg_series_equilibration_all(self._ctx, quantity, count, abs_precision, _p(e, C.c_int), _p(n, C.c_int64)))
        return e.astype(bool), n

    def mark(self):
        """Restore point (cmg_mark): what is enqueued next can be undone by rollback()."""
        self._ck(self._lib.cmg_mark(self._ctx))

    def rollback(self):
        self._ck(self._lib.cmg_rollback(self._ctx))

    def series_check(self, quantities, abs_precisions, count=None, confidence=0.95, chain=0):
        """One completion check on the device series (cmg_series_check): equilibration of every
        requested component and, if all equilibrated, the statistics of the common tail."""
        if count is None:
            count = self.n_samples
        n = len(quantities)
        q = (C.c_int * n)(*[int(v) for v in quantities])
        a = (C.c_double * n)(*[float(v) for v in abs_precisions])
        is_eq = (C.c_int * n)()
        n_eq = (C.c_int64 * n)()
        n_stats = C.c_int64()
        mean = (C.c_double * n)()
        prec = (C.c_double * n)()
        self._ck(self._lib.cmg_series_check(self._ctx, chain, n, q, a, int(count), float(confidence), is_eq, n_eq, C.byref(n_stats), mean, prec))
        return {
            "is_equilibrated": [bool(v) for v in is_eq],
            "n_equil": [int(v) for v in n_eq],
            "n_stats": int(n_stats.value),
            "mean": [float(v) for v in mean],
            "calculated_precision": [float(v) for v in prec],
        }

    # -- k-state model behind the general multi-species proposal tables --
    def kstate_set_model(self, V):
        """Switch to the k-state path: V is the symmetric K x K nearest-neighbour pair energy."""
        V = np.ascontiguousarray(V, dtype=np.float64)
        assert V.ndim == 2 and V.shape[0] == V.shape[1]
        self.n_species = int(V.shape[0])
        self._ck(self._lib.cmg_kstate_set_model(self._ctx, self.n_species, _p(V, C.c_double)))

    def kstate_set_conditions(self, temperature, mu, chain=-1):
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        assert mu.size == self.n_species
        self._ck(self._lib.cmg_kstate_set_conditions(self._ctx, int(chain), float(temperature), _p(mu, C.c_double)))

    def kstate_tables(self, chain=0):
        K, z = self.n_species, 2 * self.dim
        n = K * K * (z + 1) ** (K - 1)
        dPhi, prob = np.zeros(n), np.zeros(n)
        thr, never = np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.uint8)
        self._ck(self._lib.cmg_kstate_get_tables(self._ctx, chain, _p(dPhi, C.c_double), _p(prob, C.c_double), _p(thr, C.c_uint32), _p(never, C.c_uint8), n))
        return dPhi, prob, thr, never

    def kstate_upload(self, occ_index, chain=0):
        occ = np.ascontiguousarray(occ_index, dtype=np.int32).ravel()
        self._ck(self._lib.cmg_kstate_upload_occupation_i32(self._ctx, chain, _p(occ, C.c_int32), occ.size))

    def kstate_download(self, chain=0):
        out = np.empty(self.n_sites, dtype=np.int32)
        self._ck(self._lib.cmg_kstate_download_occupation_i32(self._ctx, chain, _p(out, C.c_int32), out.size))
        return out

    def kstate_run_passes(self, n_passes, mode=MODE_CHECKERBOARD, sample_period=0):
        self._ck(self._lib.cmg_kstate_run_passes(self._ctx, int(n_passes), int(mode), int(sample_period)))

    def kstate_samples(self, chain=0):
        """(counts [n_samples, K], bonds [n_samples, K, K]) of the sampled passes."""
        n = C.c_int64()
        self._ck(self._lib.cmg_kstate_n_samples(self._ctx, C.byref(n)))
        K = self.n_species
        counts = np.zeros((n.value, K), dtype=np.int64)
        bonds = np.zeros((n.value, K, K), dtype=np.int64)
        self._ck(self._lib.cmg_kstate_read_samples(self._ctx, chain, 0, n.value, _p(counts, C.c_int64), _p(bonds, C.c_int64)))
        return counts, bonds

    def kstate_clear_samples(self):
        self._ck(self._lib.cmg_kstate_clear_samples(self._ctx))

    # -- N-fold way driver (include/casm/monte/methods/nfold.hh) --
    def nfold_run(self, n_steps, sample_period_steps=0):
        self._ck(self._lib.cmg_nfold_run(self._ctx, int(n_steps), int(sample_period_steps)))

    def nfold_weights(self, chain=0, first=0, count=None):
        """(time increments, expected acceptance rates) of the nfold samples [first, first + count)."""
        if count is None:
            count = self.n_samples - first
        w = np.zeros(count)
        r = np.zeros(count)
        self._ck(self._lib.cmg_nfold_read_weights(self._ctx, chain, first, count, _p(w, C.c_double), _p(r, C.c_double)))
        return w, r

    def nfold_time(self, chain=0):
        t, n = C.c_double(), C.c_int64()
        self._ck(self._lib.cmg_nfold_time(self._ctx, chain, C.byref(t), C.byref(n)))
        return t.value, n.value

    # -- slab plumbing --
    def slab_half_sweep(self, colour, pass_index, sample=False):
        self._ck(self._lib.cmg_slab_half_sweep(self._ctx, int(colour), int(pass_index), int(bool(sample))))

    def slab_run_passes(self, n_passes, sample_period=0):
        """The pass loop of a slab whose neighbours are attached (fused halo push)."""
        self._ck(self._lib.cmg_slab_run_passes(self._ctx, int(n_passes), int(sample_period)))

    def slab_set_halo_exchange(self, enabled):
        self._ck(self._lib.cmg_slab_set_halo_exchange(self._ctx, int(bool(enabled))))

    def slab_boundary_ptr(self, colour, side):
        p, n = C.c_void_p(), C.c_int64()
        self._ck(self._lib.cmg_slab_boundary_ptr(self._ctx, colour, side, C.byref(p), C.byref(n)))
        return p.value, n.value

    def slab_halo_ptr(self, colour, side):
        p, n = C.c_void_p(), C.c_int64()
        self._ck(self._lib.cmg_slab_halo_ptr(self._ctx, colour, side, C.byref(p), C.byref(n)))
        return p.value, n.value

    def slab_ipc_export(self):
        buf = C.create_string_buffer(512)
        self._ck(self._lib.cmg_slab_ipc_export(self._ctx, buf, 512))
        return buf.raw

    def slab_ipc_attach(self, side, handle=None, peer=None):
        if peer is not None:
            self._ck(self._lib.cmg_slab_ipc_attach(self._ctx, side, None, 0, 1, peer._ctx))
        else:
            buf = C.create_string_buffer(handle, 512)
            self._ck(self._lib.cmg_slab_ipc_attach(self._ctx, side, buf, 512, 0, None))

    # -- introspection --
    @property
    def launch_count(self):
        n = C.c_int64()
        self._ck(self._lib.cmg_launch_count(self._ctx, C.byref(n)))
        return n.value

    @property
    def kernel_variant(self):
        return self._lib.cmg_kernel_variant(self._ctx).decode()

    def set_energy_form(self, use_nlist=True):
        """Energy form of the sampled energies (model.hh:261-285); call on an empty sample series."""
        self._ck(self._lib.cmg_set_energy_form(self._ctx, 1 if use_nlist else 0))

    def set_kernel_variant(self, name):
        self._ck(self._lib.cmg_set_kernel_variant(self._ctx, name.encode()))


# -- context-free helpers --
def host_series_stats(x, confidence=0.95, device=0):
    lib = _capi.load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    m, p, v, k = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
    check(lib.cmg_host_series_stats(device, _p(x, C.c_double), x.size, confidence, C.byref(m), C.byref(p), C.byref(v), C.byref(k)))
    return {"mean": m.value, "calculated_precision": p.value, "variance": v.value, "k_star": k.value}


def host_series_equilibration(x, abs_precision, device=0):
    lib = _capi.load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    e, n = C.c_int(), C.c_int64()
    check(lib.cmg_host_series_equilibration(device, _p(x, C.c_double), x.size, abs_precision, C.byref(e), C.byref(n)))
    return bool(e.value), n.value


def host_series_stats_weighted(x, w, confidence=0.95, method=1, n_resamples=10000, device=0):
    """BasicStatisticsCalculator()(observations, sample_weight) on the device
    (src/casm/monte/BasicStatistics.cc:144-188)."""
    lib = _capi.load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    if x.size != w.size:
        raise ValueError("Error in BasicStatisticsCalculator: observations.size() != sample_weight.size()")
    m, p, v, W, k = C.c_double(), C.c_double(), C.c_double(), C.c_double(), C.c_int64()
    check(lib.cmg_host_series_stats_weighted(device, _p(x, C.c_double), _p(w, C.c_double), x.size, confidence, method, n_resamples, C.byref(m), C.byref(p), C.byref(v), C.byref(W), C.byref(k)))
    return {"mean": m.value, "calculated_precision": p.value, "variance": v.value, "weight_sum": W.value, "k_star": k.value}


def host_series_resample(x, w, weight_sum, n_equally_spaced, device=0):
    """resample (src/casm/monte/BasicStatistics.cc:50-73) on the device."""
    lib = _capi.load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    out = np.zeros(int(n_equally_spaced), dtype=np.float64)
    check(lib.cmg_host_series_resample(device, _p(x, C.c_double), _p(w, C.c_double), x.size, float(weight_sum), out.size, _p(out, C.c_double)))
    return out


def host_series_equilibration_weighted(x, w, abs_precision, device=0):
    """Weighted branch of default_equilibration_check
    (src/casm/monte/checks/EquilibrationCheck.cc:137-161)."""
    lib = _capi.load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    e, n = C.c_int(), C.c_int64()
    check(lib.cmg_host_series_equilibration_weighted(device, _p(x, C.c_double), _p(w, C.c_double), x.size, abs_precision, C.byref(e), C.byref(n)))
    return bool(e.value), n.value


def conv_l_to_bijk(n3, n_basis, l, device=0):
    lib = _capi.load()
    n3a = (C.c_int64 * 3)(*n3)
    l = np.ascontiguousarray(l, dtype=np.int64)
    out = np.zeros((l.size, 4), dtype=np.int64)
    check(lib.cmg_conv_l_to_bijk(device, n3a, n_basis, _p(l, C.c_int64), l.size, _p(out, C.c_int64)))
    return out


def conv_bijk_to_l(n3, n_basis, bijk, device=0):
    lib = _capi.load()
    n3a = (C.c_int64 * 3)(*n3)
    b = np.ascontiguousarray(bijk, dtype=np.int64).reshape(-1, 4)
    out = np.zeros(b.shape[0], dtype=np.int64)
    check(lib.cmg_conv_bijk_to_l(device, n3a, n_basis, _p(b, C.c_int64), b.shape[0], _p(out, C.c_int64)))
    return out
