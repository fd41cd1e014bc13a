"""ctypes binding of libt4s.so (the C ABI declared in include/t4s.h).

There is no CPU fallback: if the library is missing, cannot be loaded, or the device is not
sm_100, every op raises.  `load()` is lazy so that CPU-only tooling (schema, synthetic weights,
config handling) can import the package on a box without a GPU.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("T4S_LIBRARY") or os.path.join(_HERE, "libt4s.so")   # (override: A/B builds of the diagnostics scripts)
_lock = threading.Lock()
_lib = None
_device_ok = set()

c_void_p, c_int, c_size_t, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_float


class T4sError(RuntimeError):
    pass


class MelParams(ctypes.Structure):
    _fields_ = [(n, c_int) for n in ("n_fft", "win_length", "hop", "n_mels", "preemphasis", "wav_norm", "magnitude",
                                     "out_mode", "out_dtype")]


class Operand(ctypes.Structure):
    _fields_ = [("ptr", c_void_p), ("rows", ctypes.c_int64), ("ld", ctypes.c_int64), ("nb1", ctypes.c_int64),
                ("stride1", ctypes.c_int64), ("nb2", ctypes.c_int64), ("stride2", ctypes.c_int64), ("mn_major", c_int)]


class Matrix(ctypes.Structure):
    _fields_ = [("ptr", c_void_p), ("dtype", c_int), ("ld", ctypes.c_int64), ("stride1", ctypes.c_int64),
                ("stride2", ctypes.c_int64)]


class Gemm(ctypes.Structure):
    _fields_ = [("M", c_int), ("N", c_int), ("K", c_int), ("in_dtype", c_int), ("nb1", c_int), ("nb2", c_int),
                ("split_k", c_int), ("c_split_stride", ctypes.c_int64), ("A", Operand), ("B", Operand), ("C", Matrix),
                ("aux", Matrix), ("residual", Matrix), ("bias", c_void_p), ("alpha", c_float), ("act", c_int), ("colsum", c_void_p),
                ("band_lo", c_int), ("band_hi", c_int)]


class Attn(ctypes.Structure):
    _fields_ = [("batch", c_int), ("heads", c_int), ("tokens", c_int), ("head_dim", c_int), ("scale", c_float),
                ("q", c_void_p), ("q_ld", ctypes.c_int64), ("q_bs", ctypes.c_int64),
                ("k", c_void_p), ("k_ld", ctypes.c_int64), ("k_bs", ctypes.c_int64),
                ("v", c_void_p), ("v_ld", ctypes.c_int64), ("v_bs", ctypes.c_int64),
                ("o", c_void_p), ("o_ld", ctypes.c_int64), ("o_bs", ctypes.c_int64),
                ("lse", c_void_p), ("o32", c_void_p)]


class AttnBwd(ctypes.Structure):
    _fields_ = [("fwd", Attn), ("d_o", c_void_p), ("do_ld", ctypes.c_int64), ("do_bs", ctypes.c_int64), ("delta", c_void_p),
                ("dq", c_void_p), ("dq_ld", ctypes.c_int64), ("dq_bs", ctypes.c_int64),
                ("dk", c_void_p), ("dk_ld", ctypes.c_int64), ("dk_bs", ctypes.c_int64),
                ("dv", c_void_p), ("dv_ld", ctypes.c_int64), ("dv_bs", ctypes.c_int64), ("dq32", c_void_p), ("dqkv_colsum", c_void_p)]


class RelAttn(ctypes.Structure):
    _fields_ = [("batch", c_int), ("heads", c_int), ("tokens", c_int), ("head_dim", c_int), ("scale", c_float),
                ("qu", c_void_p), ("qu_ld", ctypes.c_int64), ("qu_bs", ctypes.c_int64),
                ("qv", c_void_p), ("qv_ld", ctypes.c_int64), ("qv_bs", ctypes.c_int64),
                ("k", c_void_p), ("k_ld", ctypes.c_int64), ("k_bs", ctypes.c_int64),
                ("v", c_void_p), ("v_ld", ctypes.c_int64), ("v_bs", ctypes.c_int64),
                ("pos", c_void_p), ("pos_ld", ctypes.c_int64),
                ("o", c_void_p), ("o_ld", ctypes.c_int64), ("o_bs", ctypes.c_int64),
                ("lse", c_void_p), ("o32", c_void_p)]


class RelAttnBwd(ctypes.Structure):
    _fields_ = [("fwd", RelAttn), ("d_o", c_void_p), ("do_ld", ctypes.c_int64), ("do_bs", ctypes.c_int64), ("delta", c_void_p),
                ("dqu", c_void_p), ("dqu_ld", ctypes.c_int64), ("dqu_bs", ctypes.c_int64),
                ("dk", c_void_p), ("dk_ld", ctypes.c_int64), ("dk_bs", ctypes.c_int64),
                ("dv", c_void_p), ("dv_ld", ctypes.c_int64), ("dv_bs", ctypes.c_int64),
                ("dbd", c_void_p), ("dbd_ld", ctypes.c_int64)]


class SedLosses(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ("strong", "weak", "at", "t_strong", "t_at", "y", "yw")] + [
        ("batch", c_int), ("classes", c_int), ("strong_inner", ctypes.c_int64), ("s0", c_int), ("s1", c_int), ("w0", c_int), ("w1", c_int),
        ("w_weak", c_float), ("w_at", c_float), ("w_cons", c_float), ("w_weak_cons", c_float)]


class WindowSegment(ctypes.Structure):
    _fields_ = [("ptr", c_void_p), ("batch_stride", ctypes.c_int64), ("out_start", c_int), ("frames", c_int)]


def _declare(lib):
    P, I, Z, F, L = c_void_p, c_int, c_size_t, c_float, ctypes.c_int64
    sigs = {
        "t4s_version": (I, []),
        "t4s_last_error": (ctypes.c_char_p, []),
        "t4s_device_check": (I, []),
        "t4s_sm_count": (I, []),
        "t4s_launch_count": (ctypes.c_longlong, []),
        "t4s_adamw_step": (I, [P, P, P, P, P, Z, F, F, F, F, F, I, F, P]),
        "t4s_ema_update": (I, [P, P, P, Z, F, P]),
        "t4s_grad_pack": (I, [P, I, P, P]),
        "t4s_wav_peak": (I, [P, P, I, I, P]),
        "t4s_mel_tables_bytes": (Z, [I, I]),
        "t4s_mel_tables_init": (I, [P, P, I, I, P]),
        "t4s_mel_forward": (I, [P, P, P, P, P, P, P, I, P, I, I, I, ctypes.POINTER(MelParams), P]),
        "t4s_mel_normalize": (I, [P, P, Z, P]),
        "t4s_amp_to_db": (I, [P, P, Z, F, F, F, F, P]),
        "t4s_gemm": (I, [P, P]),   # T4sGemm* (ops.gemm packs the descriptor bytes itself)
        "t4s_reduce_splits": (I, [P, I, Z, P, I, P]),
        "t4s_attn_padded_len": (L, [I]),
        "t4s_attn_fwd": (I, [ctypes.POINTER(Attn), P]),
        "t4s_attn_bwd": (I, [ctypes.POINTER(AttnBwd), P]),
        "t4s_relattn_fwd": (I, [ctypes.POINTER(RelAttn), P]),
        "t4s_relattn_bwd": (I, [ctypes.POINTER(RelAttnBwd), P]),
        "t4s_split_tf32": (I, [ctypes.POINTER(Operand), I, P, L, I, P]),
        "t4s_layernorm_fwd": (I, [P, P, P, P, P, P, L, I, F, F, I, L, L, P]),
        "t4s_layernorm_bwd_workspace": (Z, [L, I]),
        "t4s_layernorm_bwd": (I, [P, P, P, P, P, P, P, P, P, P, P, Z, L, I, F, I, L, L, P]),
        "t4s_colsum_workspace": (Z, [L, I]),
        "t4s_colsum": (I, [P, I, L, I, L, P, Z, P, I, P]),
        "t4s_gelu_bwd": (I, [P, P, P, Z, I, P]),
        "t4s_softmax_fwd": (I, [P, P, L, I, L, L, I, P]),
        "t4s_softmax_bwd": (I, [P, P, L, I, L, L, I, P]),
        "t4s_relpos_softmax_fwd": (I, [P, P, P, L, I, L, L, L, I, P]),
        "t4s_relpos_softmax_bwd": (I, [P, P, P, L, I, L, L, L, I, P]),
        "t4s_patch_im2col": (I, [P, I, P, I, I, I, I, I, I, I, I, P]),
        "t4s_add2": (I, [P, L, P, L, P, L, L, I, F, F, I, P]),
        "t4s_patch_im2col_windows": (I, [P, I, P, I, I, I, I, ctypes.POINTER(c_int), I, I, I, I, I, P]),
        "t4s_window_overlap_add_fwd": (I, [ctypes.POINTER(WindowSegment), I, P, I, I, I, I, P]),
        "t4s_window_overlap_add_bwd": (I, [P, ctypes.POINTER(WindowSegment), I, I, I, I, I, P]),
        "t4s_mask_rows_fwd": (I, [P, P, P, P, P, L, I, I, P]),
        "t4s_mask_rows_bwd_workspace": (Z, [L, I]),
        "t4s_mask_rows_bwd": (I, [P, P, P, P, I, P, P, P, Z, L, I, I, P]),
        "t4s_im2col3x3": (I, [P, I, L, L, L, P, I, I, I, I, I, I, P]),
        "t4s_col2im3x3": (I, [P, P, I, I, I, I, I, I, P]),
        "t4s_chan_stats_workspace": (Z, [L, I]),
        "t4s_batchnorm_fwd": (I, [P, P, I, L, I, P, P, P, P, F, F, I, P, P, P, Z, P]),
        "t4s_batchnorm_bwd": (I, [P, P, I, L, I, P, P, P, I, P, P, P, P, Z, P]),
        "t4s_gate_fwd": (I, [P, P, P, Z, F, ctypes.c_uint64, I, P]),
        "t4s_gate_bwd": (I, [P, P, P, P, P, Z, F, ctypes.c_uint64, I, P]),
        "t4s_avgpool_fwd": (I, [P, P, I, I, I, I, I, I, I, P]),
        "t4s_avgpool_bwd": (I, [P, P, I, I, I, I, I, I, I, P]),
        "t4s_scale_add_fwd": (I, [P, P, P, P, Z, I, P]),
        "t4s_scale_add_bwd": (I, [P, P, P, P, P, P, Z, I, P]),
        "t4s_l2norm_fwd": (I, [P, P, P, L, I, I, P]),
        "t4s_l2norm_bwd": (I, [P, P, P, P, L, I, I, P]),
        "t4s_proto_act_fwd": (I, [P, P, Z, F, F, P]),
        "t4s_proto_act_bwd": (I, [P, P, P, P, Z, F, F, P]),
        "t4s_query_pool_fwd": (I, [P, P, P, F, P, P, I, I, I, P]),
        "t4s_query_pool_bwd": (I, [P, P, P, P, P, P, F, P, P, I, I, I, P]),
        "t4s_mask_scores": (I, [P, P, L, I, L, I, I, P]),
        "t4s_dropout": (I, [P, P, Z, F, ctypes.c_uint64, I, P]),
        "t4s_roll_rows": (I, [P, P, P, I, I, I, P]),
        "t4s_mixup": (I, [P, P, P, I, L, F, F, I, P]),
        "t4s_median_filter": (I, [P, P, ctypes.POINTER(c_int), I, I, I, P]),
        "t4s_freq_warp": (I, [P, P, P, P, I, I, I, P]),
        "t4s_scaler_instance": (I, [P, P, I, L, I, F, P]),
        "t4s_scaler_dataset": (I, [P, P, L, P, P, I, F, P]),
        "t4s_rank_filter": (I, [P, P, ctypes.POINTER(c_int), I, I, I, I, P]),
        "t4s_event_sweep": (I, [P, P, P, I, I, I, I, P, P, P, P]),
        "t4s_sed_losses_fwd": (I, [ctypes.POINTER(SedLosses), P, P, P]),
        "t4s_sed_losses_bwd": (I, [ctypes.POINTER(SedLosses), P, P, P, P, P]),
        "t4s_add_rowbias": (I, [P, P, P, L, I, P]),
        "t4s_patch_posbias": (I, [P, P, P, I, I, I, I, I, P]),
        "t4s_cls_dist_tokens": (I, [P, I, P, P, P, I, L, I, P]),
        "t4s_patch_small_grads": (I, [P, I, P, P, P, P, P, P, P, I, L, I, I, I, I, I, P]),
        "t4s_fpool_mean_fwd": (I, [P, P, I, I, I, I, I, P]),
        "t4s_fpool_mean_bwd": (I, [P, P, I, I, I, I, I, P]),
        "t4s_pad_interp_fwd": (I, [P, P, I, I, I, I, I, I, P]),
        "t4s_pad_interp_bwd": (I, [P, P, I, I, I, I, I, I, P]),
        "t4s_add_rowvec": (I, [P, L, P, P, L, I, F, I, P]),
        "t4s_convert": (I, [P, I, P, I, Z, P]),
        "t4s_sed_pool_fwd": (I, [P, P, F, P, P, I, I, I, P]),
        "t4s_sed_pool_bwd": (I, [P, P, P, P, F, P, I, I, I, P]),
        "t4s_sigmoid_fwd": (I, [P, P, Z, P]),
        "t4s_sigmoid_bwd": (I, [P, P, P, Z, P]),
        "t4s_bce_fwd": (I, [P, P, Z, P, P, P]),
        "t4s_bce_bwd": (I, [P, P, P, Z, P, P]),
        "t4s_mse_fwd": (I, [P, P, P, L, I, I, P, P, P]),
        "t4s_mse_bwd": (I, [P, P, P, L, I, I, P, P, P, P, P]),
        "t4s_attnpool_fwd": (I, [P, P, P, P, I, I, I, I, L, I, P]),
        "t4s_attnpool_bwd": (I, [P, P, P, P, P, P, I, I, I, I, L, I, P]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return sigs


def load():
    """Return the loaded library, building it first if the sources are newer and nvcc is present."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            try:
                from . import build as _build
                _build.build()
            except Exception as e:  # noqa: BLE001
                raise T4sError(f"libt4s.so is missing at {LIB_PATH} and could not be built ({e}); "
                               "run `python -m transformer4sed_b200.build`.  There is no CPU fallback.") from e
        try:
            lib = ctypes.CDLL(LIB_PATH)
        except OSError as e:
            raise T4sError(f"cannot load {LIB_PATH}: {e}.  There is no CPU fallback.") from e
        _declare(lib)
        _lib = lib
    return _lib


def exported_symbols():
    """Names declared in include/t4s.h, parsed from the header (used by the CPU ABI test)."""
    import re
    hdr = open(os.path.join(os.path.dirname(_HERE), "include", "t4s.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(t4s_[a-z0-9_]+)\s*\(", hdr)))


class LaunchProfiler:
    """Brackets every C-ABI call with CUDA events on the launching stream (bench.py's instrumented step).
    `records` = [(name, key, start_event, end_event)]; keys carry GEMM shapes so time can be grouped per op."""

    def __init__(self):
        self.records = []

    def timed(self, name, key, fn):
        import torch
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = fn()
        e.record()
        self.records.append((name, key, s, e))
        return rc

    def summary(self):
        agg = {}
        for name, key, s, e in self.records:
            k = name if key is None else f"{name}{key}"
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += s.elapsed_time(e)
        return dict(sorted(agg.items(), key=lambda kv: -kv[1][1]))


profiler = None  # set to a LaunchProfiler to time every launch (never inside a timed benchmark region)


def check(rc, what=""):
    if rc != 0:
        msg = load().t4s_last_error().decode(errors="replace")
        raise T4sError(f"{what} failed with code {rc}: {msg}")


def ensure_device(t):
    """Validate that tensor `t` lives on an sm_100 CUDA device and make it current."""
    import torch
    if not t.is_cuda:
        raise T4sError("transformer4sed_b200 ops need CUDA tensors on a B200 (sm_100a); there is no CPU fallback")
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if idx not in _device_ok:
        with torch.cuda.device(idx):
            check(load().t4s_device_check(), "t4s_device_check")
        _device_ok.add(idx)
    return idx


def stream_ptr():
    """Raw handle of the current stream of the current device (the C entry points torch itself uses: `torch.cuda.current_stream()`
    costs several microseconds of Python per call, and it is called once per kernel launch)."""
    import torch
    return c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


class _NullGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NULL_GUARD = _NullGuard()


def device_guard(device):
    """`with device_guard(t.device):` = `with torch.cuda.device(t.device):` without the context-manager cost when that device is
    already current (the normal case: one process per GPU)."""
    import torch
    idx = device.index if isinstance(device, torch.device) else device
    if idx is None or idx == torch._C._cuda_getDevice():
        return _NULL_GUARD
    return torch.cuda.device(idx)


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)
