"""ctypes binding of the C ABI (include/casm_monte_gpu.h).

The shared library is built in-tree by ``casmcode_monte_b200.build`` (nvcc,
sm_100a).  There is no CPU fallback: if the library is missing, or no CUDA
device is present, the calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcasm_monte_b200.so")

CMG_OK = 0
MODE_CHECKERBOARD = 0
MODE_SERIAL_REFERENCE = 1
Q_PARAM_COMPOSITION = 0
Q_FORMATION_ENERGY = 1
Q_POTENTIAL_ENERGY = 2
KB = 8.6173303e-05

_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)
_f64p = C.POINTER(C.c_double)
_intp = C.POINTER(C.c_int)
_ctx = C.c_void_p

# name -> (restype, argtypes); every symbol include/casm_monte_gpu.h declares
SIGNATURES = {
    "cmg_abi_version": (C.c_int, []),
    "cmg_last_global_error": (C.c_char_p, []),
    "cmg_last_error": (C.c_char_p, [_ctx]),
    "cmg_device_count": (C.c_int, [_intp]),
    "cmg_create": (C.c_int, [C.c_int, _i64p, C.c_int, C.c_int, C.POINTER(_ctx)]),
    "cmg_destroy": (C.c_int, [_ctx]),
    "cmg_set_stream": (C.c_int, [_ctx, C.c_void_p]),
    "cmg_sync": (C.c_int, [_ctx]),
    "cmg_n_sites": (C.c_int, [_ctx, _i64p]),
    "cmg_create_slab": (C.c_int, [C.c_int, _i64p, C.c_int64, C.c_int64, C.c_int, C.POINTER(_ctx)]),
    "cmg_slab_boundary_ptr": (C.c_int, [_ctx, C.c_int, C.c_int, C.POINTER(C.c_void_p), _i64p]),
    "cmg_slab_halo_ptr": (C.c_int, [_ctx, C.c_int, C.c_int, C.POINTER(C.c_void_p), _i64p]),
    "cmg_slab_ipc_export": (C.c_int, [_ctx, C.c_void_p, C.c_int64]),
    "cmg_slab_ipc_attach": (C.c_int, [_ctx, C.c_int, C.c_void_p, C.c_int64, C.c_int, _ctx]),
    "cmg_slab_half_sweep": (C.c_int, [_ctx, C.c_int, C.c_uint64, C.c_int]),
    "cmg_slab_run_passes": (C.c_int, [_ctx, C.c_int64, C.c_int64]),
    "cmg_slab_set_halo_exchange": (C.c_int, [_ctx, C.c_int]),
    "cmg_set_model": (C.c_int, [_ctx, C.c_double, C.c_int]),
    "cmg_set_conditions": (C.c_int, [_ctx, C.c_int, C.c_double, C.c_double]),
    "cmg_get_tables": (C.c_int, [_ctx, C.c_int, _f64p, _f64p, _u32p]),
    "cmg_upload_occupation_i32": (C.c_int, [_ctx, C.c_int, _i32p, C.c_int64]),
    "cmg_download_occupation_i32": (C.c_int, [_ctx, C.c_int, _i32p, C.c_int64]),
    "cmg_upload_occupation_i32_dev": (C.c_int, [_ctx, C.c_int, C.c_void_p, C.c_int64]),
    "cmg_download_occupation_i32_dev": (C.c_int, [_ctx, C.c_int, C.c_void_p, C.c_int64]),
    "cmg_upload_occupation_i8": (C.c_int, [_ctx, C.c_int, C.POINTER(C.c_int8), C.c_int64]),
    "cmg_download_occupation_i8": (C.c_int, [_ctx, C.c_int, C.POINTER(C.c_int8), C.c_int64]),
    "cmg_upload_occupation_bits": (C.c_int, [_ctx, C.c_int, _u8p, C.c_int64]),
    "cmg_download_occupation_bits": (C.c_int, [_ctx, C.c_int, _u8p, C.c_int64]),
    "cmg_fill_occupation": (C.c_int, [_ctx, C.c_int, C.c_int]),
    "cmg_get_occ": (C.c_int, [_ctx, C.c_int, C.c_int64, _i32p]),
    "cmg_set_occ": (C.c_int, [_ctx, C.c_int, C.c_int64, C.c_int32]),
    "cmg_event_delta": (C.c_int, [_ctx, C.c_int, C.c_int, _i64p, _i32p, _f64p, _f64p]),
    "cmg_randomize_occupation": (C.c_int, [_ctx, C.c_int, C.c_uint64, C.c_double]),
    "cmg_seed_philox": (C.c_int, [_ctx, C.c_uint64]),
    "cmg_set_philox_rounds": (C.c_int, [_ctx, C.c_int]),
    "cmg_set_pass_counter": (C.c_int, [_ctx, C.c_uint64]),
    "cmg_set_chain_offset": (C.c_int, [_ctx, C.c_int64]),
    "cmg_seed_mt19937_64": (C.c_int, [_ctx, C.c_int, C.c_uint64]),
    "cmg_set_mt19937_64_state": (C.c_int, [_ctx, C.c_int, _u64p, C.c_int]),
    "cmg_get_mt19937_64_state": (C.c_int, [_ctx, C.c_int, _u64p, _intp]),
    "cmg_rng_draw": (C.c_int, [_ctx, C.c_int, C.c_int, _i64p, _f64p, _u8p, _i64p, _f64p]),
    "cmg_run_passes": (C.c_int, [_ctx, C.c_int64, C.c_int, C.c_int64]),
    "cmg_counters": (C.c_int, [_ctx, C.c_int, _i64p, _i64p, _i64p]),
    "cmg_reset_counters": (C.c_int, [_ctx]),
    "cmg_sample_now": (C.c_int, [_ctx, C.c_int, _i64p, _i64p]),
    "cmg_line_dots": (C.c_int, [_ctx, C.c_int, _i64p, _i64p]),
    "cmg_n_samples": (C.c_int, [_ctx, _i64p]),
    "cmg_clear_samples": (C.c_int, [_ctx]),
    "cmg_read_samples_sb": (C.c_int, [_ctx, C.c_int, C.c_int64, C.c_int64, _i64p, _i64p]),
    "cmg_read_samples": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int64, C.c_int64, _f64p]),
    "cmg_delta_e_probe": (C.c_int, [_ctx, C.c_int, _f64p]),
    "cmg_accept_probe": (C.c_int, [_ctx, C.c_int, _f64p, _u8p]),
    "cmg_series_stats": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_double, _f64p, _f64p, _f64p, _i64p]),
    "cmg_series_equilibration": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int64, C.c_double, _intp, _i64p]),
    "cmg_series_stats_all": (C.c_int, [_ctx, C.c_int, _i64p, C.c_int64, C.c_double, _f64p, _f64p, _f64p, _i64p]),
    "cmg_series_equilibration_all": (C.c_int, [_ctx, C.c_int, C.c_int64, C.c_double, _intp, _i64p]),
    "cmg_mark": (C.c_int, [_ctx]),
    "cmg_rollback": (C.c_int, [_ctx]),
    "cmg_series_check": (C.c_int, [_ctx, C.c_int, C.c_int, _intp, _f64p, C.c_int64, C.c_double, _intp, _i64p, _i64p, _f64p, _f64p]),
    "cmg_series_check_prefetch": (C.c_int, [_ctx, C.c_int, C.c_int, _intp, _f64p, C.c_int64, C.c_double]),
    "cmg_host_series_stats": (C.c_int, [C.c_int, _f64p, C.c_int64, C.c_double, _f64p, _f64p, _f64p, _i64p]),
    "cmg_host_series_equilibration": (C.c_int, [C.c_int, _f64p, C.c_int64, C.c_double, _intp, _i64p]),
    "cmg_host_series_stats_weighted": (C.c_int, [C.c_int, _f64p, _f64p, C.c_int64, C.c_double, C.c_int, C.c_int64, _f64p, _f64p, _f64p, _f64p, _i64p]),
    "cmg_host_series_resample": (C.c_int, [C.c_int, _f64p, _f64p, C.c_int64, C.c_double, C.c_int64, _f64p]),
    "cmg_host_series_equilibration_weighted": (C.c_int, [C.c_int, _f64p, _f64p, C.c_int64, C.c_double, _intp, _i64p]),
    "cmg_conv_l_to_bijk": (C.c_int, [C.c_int, _i64p, C.c_int64, _i64p, C.c_int64, _i64p]),
    "cmg_conv_bijk_to_l": (C.c_int, [C.c_int, _i64p, C.c_int64, _i64p, C.c_int64, _i64p]),
    "cmg_conv_general_l_to_bijk": (C.c_int, [C.c_int, _i64p, C.c_int64, _i64p, C.c_int64, _i64p]),
    "cmg_conv_general_bijk_to_l": (C.c_int, [C.c_int, _i64p, C.c_int64, _i64p, C.c_int64, _i64p]),
    "cmg_kstate_set_model": (C.c_int, [_ctx, C.c_int, _f64p]),
    "cmg_kstate_set_conditions": (C.c_int, [_ctx, C.c_int, C.c_double, _f64p]),
    "cmg_kstate_get_tables": (C.c_int, [_ctx, C.c_int, _f64p, _f64p, _u32p, _u8p, C.c_int64]),
    "cmg_kstate_upload_occupation_i32": (C.c_int, [_ctx, C.c_int, _i32p, C.c_int64]),
    "cmg_kstate_download_occupation_i32": (C.c_int, [_ctx, C.c_int, _i32p, C.c_int64]),
    "cmg_kstate_run_passes": (C.c_int, [_ctx, C.c_int64, C.c_int, C.c_int64]),
    "cmg_kstate_n_samples": (C.c_int, [_ctx, _i64p]),
    "cmg_kstate_clear_samples": (C.c_int, [_ctx]),
    "cmg_kstate_read_samples": (C.c_int, [_ctx, C.c_int, C.c_int64, C.c_int64, _i64p, _i64p]),
    "cmg_nfold_run": (C.c_int, [_ctx, C.c_int64, C.c_int64]),
    "cmg_nfold_read_weights": (C.c_int, [_ctx, C.c_int, C.c_int64, C.c_int64, _f64p, _f64p]),
    "cmg_nfold_time": (C.c_int, [_ctx, C.c_int, _f64p, _i64p]),
    "cmg_set_energy_form": (C.c_int, [_ctx, C.c_int]),
    "cmg_launch_count": (C.c_int, [_ctx, _i64p]),
    "cmg_kernel_variant": (C.c_char_p, [_ctx]),
    "cmg_set_kernel_variant": (C.c_int, [_ctx, C.c_char_p]),
}

_lib = None


class CmgError(RuntimeError):
    """A C-ABI call failed (maps the reference's std::runtime_error)."""

    def __init__(self, code, message):
        super().__init__(f"[cmg {code}] {message}")
        self.code = code


def load():
    """Load libcasm_monte_b200.so and attach signatures.  Raises if missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("CMG_LIB_PATH", LIB_PATH)  # developer override: A/B builds of the same ABI
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback."
        )
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, ctx=None):
    if rc != CMG_OK:
        lib = load()
        msg = lib.cmg_last_error(ctx) if ctx else lib.cmg_last_global_error()
        raise CmgError(rc, (msg or b"").decode())
