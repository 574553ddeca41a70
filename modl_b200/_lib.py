"""ctypes binding of libmodl_b200.so (the C ABI declared in include/modl_b200.h).

The library is built in-tree by `python -m modl_b200.build` (or __graft_entry__.build()).
There is no CPU fallback: if the shared object is missing this module raises, and every
device entry point returns an error code that is turned into an exception here.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MODL_B200_LIB") or os.path.join(_HERE, "libmodl_b200.so")     # the override is for A/B builds

MODL_OK, MODL_EINVAL, MODL_ECUDA, MODL_ENOTSPD, MODL_ENOMEM = 0, 1, 2, 3, 4
AGG = {"masked": 0, "full": 1, "average": 2}

vp, i64, u64, f64, f32, ci = C.c_void_p, C.c_int64, C.c_uint64, C.c_double, C.c_float, C.c_int


class StepParams(C.Structure):
    """struct modl_step_params (include/modl_b200.h)."""
    _fields_ = [
        ("n_samples", i64), ("n_features", i64), ("n_components", i64), ("batch_size", i64),
        ("X", vp), ("ldx", i64), ("indices", vp), ("h_subset", vp), ("subset_len", i64),
        ("h_order", vp), ("w_sample", vp), ("w", f64),
        ("components", vp), ("code", vp), ("C", vp), ("B", vp), ("comp_norm", vp),
        ("G_full", vp), ("Dx_average", vp), ("G_average", vp),
        ("reduction", f64), ("code_alpha", f64), ("code_l1_ratio", f64), ("comp_l1_ratio", f64),
        ("tol", f64), ("step_size", f64),
        ("max_iter", ci), ("code_pos", ci), ("comp_pos", ci), ("Dx_agg", ci), ("G_agg", ci),
        ("optimizer_sgd", ci),
        ("sweeps", vp),
        ("phases", ci), ("global_batch", i64), ("stats_inc", vp), ("inc_sub", vp), ("ev_after_apply_sub", vp),
        ("slot", ci), ("sm_avail", ci), ("start_flag", vp), ("start_serial", C.c_uint32), ("h_inputs_mapped", ci),
    ]

class FitParams(C.Structure):
    """struct modl_fit_params (include/modl_b200.h)."""
    _fields_ = [
        ("n_samples", i64), ("n_features", i64), ("n_components", i64), ("batch_size", i64),
        ("components", vp), ("code", vp), ("C", vp), ("B", vp), ("comp_norm", vp),
        ("G_full", vp), ("Dx_average", vp), ("G_average", vp),
        ("reduction", f64), ("learning_rate", f64), ("code_alpha", f64), ("code_l1_ratio", f64),
        ("comp_l1_ratio", f64), ("tol", f64), ("step_size", f64),
        ("max_iter", ci), ("code_pos", ci), ("comp_pos", ci), ("Dx_agg", ci), ("G_agg", ci), ("optimizer_sgd", ci),
    ]


class FitBatches(C.Structure):
    """struct modl_fit_batches (include/modl_b200.h)."""
    _fields_ = [
        ("X", vp), ("ldx", i64), ("n_rows", i64), ("x_location", ci),
        ("h_sample_indices", vp), ("sampler", vp), ("h_n_iter", vp), ("h_sample_n_iter", vp),
        ("update_counters", ci), ("h_orders", vp), ("h_w_sample", vp), ("sweeps", vp),
        ("h_last_subset", vp), ("h_last_subset_len", vp), ("h_code_out", vp), ("wait_host", ci),
        ("fence", ci),
    ]


PHASE_CODE, PHASE_STATS, PHASE_APPLY, PHASE_DICT, PHASE_APPLY_SUB, PHASE_APPLY_B = 1, 2, 4, 8, 16, 32
PHASE_STATS_SUB, PHASE_STATS_B, PHASE_REUSE_SUBSET = 64, 128, 256
PHASE_PREFETCH, PHASE_INPUTS_READY, PHASE_FUSED_APPLY = 512, 1024, 2048


class ModlError(RuntimeError):
    def __init__(self, status, message):
        RuntimeError.__init__(self, "modl_b200 error %d: %s" % (status, message))
        self.status = status


def _declare(L):
    def fn(name, res, args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
        return f

    fn("modl_version", ci, [])
    fn("modl_struct_sizes", None, [vp])
    fn("modl_last_error", C.c_char_p, [])
    fn("modl_rs_create", vp, [u64])
    fn("modl_rs_destroy", None, [vp])
    fn("modl_rs_seed", None, [vp, u64])
    fn("modl_rs_randint", i64, [vp, u64])
    fn("modl_rs_binomial", i64, [vp, i64, f64])
    fn("modl_rs_permutation", None, [vp, vp, i64])
    fn("modl_rs_shuffle", None, [vp, vp, i64])
    fn("modl_rs_shuffle_with_trace", None, [vp, i64, vp, vp])
    fn("modl_sampler_create", vp, [i64, ci, ci, u64])
    fn("modl_sampler_destroy", None, [vp])
    fn("modl_sampler_yield_subset", i64, [vp, f64, vp])
    fn("modl_batch_weight", f64, [i64, i64, f64, f64])
    fn("modl_ctx_create", ci, [ci, C.POINTER(vp)])
    fn("modl_ctx_destroy", None, [vp])
    fn("modl_ctx_sm_count", ci, [vp])
    fn("modl_ctx_launch_count", i64, [vp])
    fn("modl_ctx_set_option", ci, [vp, C.c_char_p, ci])
    fn("modl_ctx_check_info", ci, [vp, vp])
    fn("modl_ctx_profile", ci, [vp, ci])
    fn("modl_ctx_profile_read", ci, [vp, vp, vp])
    fn("modl_fit_create", ci, [vp, C.POINTER(vp)])
    fn("modl_fit_destroy", None, [vp])
    fn("modl_fit_set_option", ci, [vp, C.c_char_p, ci])
    fn("modl_fit_synchronize", ci, [vp])
    fn("modl_fit_graph_stats", ci, [vp, vp])
    fn("modl_fit_trace", ci, [vp, ci])
    fn("modl_fit_trace_read", ci, [vp, vp, ci, vp])
    fn("modl_nccl_unique_id", ci, [vp, i64])
    fn("modl_fit_set_comm", ci, [vp, ci, ci, vp, vp])
    for sfx, real in (("f32", f32), ("f64", f64)):
        fn("modl_partial_fit_" + sfx, ci, [vp, C.POINTER(FitParams), C.POINTER(FitBatches), vp])
        fn("modl_enet_norm_" + sfx, ci, [vp, vp, i64, i64, i64, real, vp, vp])
        fn("modl_enet_projection_" + sfx, ci, [vp, vp, vp, i64, i64, i64, vp, real, vp])
        fn("modl_enet_scale_" + sfx, ci, [vp, vp, i64, i64, i64, real, real, vp])
        fn("modl_gram_dx_" + sfx, ci, [vp, vp, i64, vp, i64, vp, i64, i64, i64, i64, real, vp, vp, vp, vp])
        for nm in ("modl_enet_regression_single_gram_", "modl_enet_regression_multi_gram_"):
            fn(nm + sfx, ci, [vp, vp, vp, vp, i64, i64, vp, vp, vp, i64, i64, real, real, ci, real, ci, vp, vp])
        fn("modl_update_G_average_" + sfx, ci, [vp, vp, vp, vp, vp, i64, i64, vp])
        fn("modl_update_Dx_average_" + sfx, ci, [vp, vp, vp, vp, vp, i64, i64, vp])
        fn("modl_update_stats_" + sfx, ci, [vp, vp, vp, vp, i64, vp, vp, i64, f64, i64, i64, i64, ci, vp])
        fn("modl_update_dict_" + sfx, ci, [vp, vp, i64, vp, i64, vp, vp, vp, vp, i64, vp, i64, i64,
                                           real, ci, ci, f64, f64, vp])
        fn("modl_batch_fit_" + sfx, ci, [vp, C.POINTER(StepParams), vp])
        fn("modl_recsys_gram_dx_" + sfx, ci, [vp, vp, i64, vp, vp, vp, vp, i64, i64, i64, i64, f64, vp, vp, vp])
        fn("modl_recsys_update_B_" + sfx, ci, [vp, vp, i64, vp, vp, vp, vp, vp, vp, i64, i64, f64, i64, vp])
        fn("modl_recsys_update_C_" + sfx, ci, [vp, vp, vp, vp, i64, i64, f64, vp])
        fn("modl_recsys_sync_transposed_" + sfx, ci, [vp, vp, i64, vp, i64, vp, i64, i64, vp])
        fn("modl_recsys_predict_" + sfx, ci, [vp, vp, vp, i64, vp, vp, i64, i64, vp, vp])


# every symbol include/modl_b200.h declares (tests check the library exports them all)
EXPORTED = (
    ["modl_version", "modl_struct_sizes", "modl_last_error", "modl_rs_create", "modl_rs_destroy", "modl_rs_seed",
     "modl_rs_randint", "modl_rs_binomial", "modl_rs_permutation", "modl_rs_shuffle",
     "modl_rs_shuffle_with_trace", "modl_sampler_create", "modl_sampler_destroy",
     "modl_sampler_yield_subset", "modl_batch_weight", "modl_ctx_create", "modl_ctx_destroy",
     "modl_ctx_sm_count", "modl_ctx_launch_count", "modl_ctx_set_option", "modl_ctx_check_info",
     "modl_ctx_profile", "modl_ctx_profile_read", "modl_fit_create", "modl_fit_destroy", "modl_fit_set_option",
     "modl_fit_synchronize", "modl_fit_graph_stats", "modl_fit_trace", "modl_fit_trace_read", "modl_nccl_unique_id", "modl_fit_set_comm"]
    + [n + s for s in ("f32", "f64") for n in (
        "modl_enet_norm_", "modl_enet_projection_", "modl_enet_scale_", "modl_gram_dx_",
        "modl_enet_regression_single_gram_", "modl_enet_regression_multi_gram_",
        "modl_update_G_average_", "modl_update_Dx_average_", "modl_update_stats_",
        "modl_update_dict_", "modl_batch_fit_", "modl_partial_fit_", "modl_recsys_gram_dx_", "modl_recsys_update_B_",
        "modl_recsys_update_C_", "modl_recsys_sync_transposed_", "modl_recsys_predict_")])

_lib = None


def lib():
    """Load the shared library (once).  Raises ImportError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "modl_b200: %s is missing -- build the CUDA extension first with "
                "`python -m modl_b200.build` (there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        _declare(L)
        _lib = L
    return _lib


def check(status):
    if status != MODL_OK:
        raise ModlError(status, lib().modl_last_error().decode("utf-8", "replace"))


_contexts = {}


class Context(object):
    """One modl_ctx per CUDA device (workspace arena + launch geometry)."""

    def __init__(self, device_index):
        h = vp()
        check(lib().modl_ctx_create(int(device_index), C.byref(h)))
        self.handle = h
        self.device_index = int(device_index)

    @property
    def sm_count(self):
        return lib().modl_ctx_sm_count(self.handle)

    @property
    def launch_count(self):
        return int(lib().modl_ctx_launch_count(self.handle))

    def set_option(self, name, value):
        check(lib().modl_ctx_set_option(self.handle, name.encode(), int(value)))

    PHASES = ("gather", "gram", "average", "code", "stats", "dict_prep", "dict_bcd", "dict_post")

    def profile(self, enable):
        check(lib().modl_ctx_profile(self.handle, int(bool(enable))))

    def profile_read(self):
        """-> ({phase: total_ms}, steps) accumulated since profile(True)."""
        ms = (C.c_double * len(self.PHASES))()
        steps = C.c_int64(0)
        check(lib().modl_ctx_profile_read(self.handle, ms, C.byref(steps)))
        return dict(zip(self.PHASES, list(ms))), int(steps.value)

    def check_info(self, stream):
        check(lib().modl_ctx_check_info(self.handle, vp(stream)))


class FitLoop(object):
    """One modl_fit per estimator: the streams, events, staging slots and (sharded) NCCL communicators of its
    minibatch loop (modl_partial_fit_*)."""

    def __init__(self, ctx):
        h = vp()
        check(lib().modl_fit_create(ctx.handle, C.byref(h)))
        self.handle = h
        self.ctx = ctx
        self.world = 1

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                lib().modl_fit_destroy(h)
            except Exception:
                pass

    def set_option(self, name, value):
        check(lib().modl_fit_set_option(self.handle, name.encode(), int(value)))

    def synchronize(self):
        check(lib().modl_fit_synchronize(self.handle))

    def graph_stats(self):
        """-> {graph launches, executable graphs rebuilt (topology changed), calls that fell back to plain launches}."""
        out = (C.c_int64 * 4)()
        check(lib().modl_fit_graph_stats(self.handle, out))
        return {"launches": int(out[0]), "rebuilt": int(out[1]), "fallbacks": int(out[2]), "gate_timeouts": int(out[3])}

    TRACE_POINTS = ("h2d_start", "h2d_end", "prefetch_start", "prefetch_end", "main_start", "codes_done", "dict_done",
                    "stats_b_start", "stats_b_end", "d2h_done")

    def trace(self, steps):
        """Arm a timeline of the next `steps` steps (0 = off); read it back with trace_read()."""
        check(lib().modl_fit_trace(self.handle, int(steps)))
        self._trace_cap = int(steps)

    def trace_read(self):
        """-> list of {point: ms since arming} for the traced steps."""
        cap = getattr(self, "_trace_cap", 0)
        buf = (C.c_double * (cap * len(self.TRACE_POINTS)))()
        n = C.c_int(0)
        check(lib().modl_fit_trace_read(self.handle, buf, cap, C.byref(n)))
        out = []
        for i in range(n.value):
            row = buf[i * len(self.TRACE_POINTS):(i + 1) * len(self.TRACE_POINTS)]
            out.append({name: v for name, v in zip(self.TRACE_POINTS, row) if v == v})
        return out

    def set_comm(self, world, rank, id_main, id_side):
        check(lib().modl_fit_set_comm(self.handle, int(world), int(rank), id_main, id_side))
        self.world = int(world)


def nccl_unique_id():
    """128 bytes identifying a new NCCL communicator (drawn on one rank, shipped to the others)."""
    buf = C.create_string_buffer(128)
    check(lib().modl_nccl_unique_id(buf, 128))
    return buf.raw


def get_context(device_index):
    ctx = _contexts.get(device_index)
    if ctx is None:
        ctx = _contexts[device_index] = Context(device_index)
    return ctx


def sfx_of(dtype):
    import torch
    if dtype == torch.float32:
        return "f32"
    if dtype == torch.float64:
        return "f64"
    raise TypeError("modl_b200 supports float32 and float64, got %s" % (dtype,))
