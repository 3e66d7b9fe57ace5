"""ctypes binding of libvangan_b200.so (the C ABI declared in include/vangan_b200.h).

torch is used only for device memory (`tensor.data_ptr()`) and the current CUDA stream.  There is
NO fallback: if the shared library is missing, fails to load, or a call returns an error code, an
exception is raised.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvangan_b200.so")

VG_F32, VG_BF16 = 0, 1
IN_RELU_INPUT = 0x100
IN_BATCH_STATS = 0x200
IN_DY_SCRATCH = 0x400
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_TANH = 0, 1, 2, 3
PAD_ZERO, PAD_REFLECT = 0, 1
_ERR = {-1: "invalid argument", -2: "unsupported shape", -3: "workspace too small", -4: "CUDA error"}


class VgError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [("N", C.c_int), ("ID", C.c_int), ("IH", C.c_int), ("IW", C.c_int), ("Cin", C.c_int), ("Cout", C.c_int),
                ("K", C.c_int), ("stride", C.c_int), ("x_dtype", C.c_int), ("y_dtype", C.c_int), ("act", C.c_int),
                ("dx_lo", C.c_int), ("dx_hi", C.c_int)]


class InDesc(C.Structure):
    _fields_ = [("N", C.c_int), ("D", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int), ("dtype", C.c_int),
                ("act", C.c_int), ("slope", C.c_float), ("pad_lo", C.c_int), ("pad_hi", C.c_int), ("pad_mode", C.c_int),
                ("noise_std", C.c_float), ("seed", C.c_ulonglong), ("seed_dev", C.c_void_p)]


_P, _I, _F, _Z, _LL, _ULL = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_longlong, C.c_ulonglong
_CD, _ID = C.POINTER(ConvDesc), C.POINTER(InDesc)

# name -> (restype, argtypes); must list every symbol of include/vangan_b200.h (tests/test_abi.py checks)
SIGNATURES = {
    "vg_abi_version": (_I, []),
    "vg_launch_count": (_ULL, []),
    "vg_tc_launch_count": (_ULL, []),
    "vg_conv3d_packed_bytes": (_Z, [_CD, _I]),
    "vg_conv3d_pack_weights": (_I, [_CD, _P, _P, _P, _P]),
    "vg_conv3d_pack_jobs": (_I, [_CD, _P, _P, _P, _P, _I]),
    "vg_pack_job_bytes": (_Z, []),
    "vg_pack_job_total": (_LL, [_P, _I]),
    "vg_pack_run": (_I, [_P, _P, _I, _LL, _P]),
    "vg_conv3d_fwd": (_I, [_CD, _P, _P, _P, _P, _P]),
    "vg_conv3d_dgrad": (_I, [_CD, _P, _P, _P, _P]),
    "vg_conv3d_wgrad": (_I, [_CD, _P, _P, _P, _P, _P]),
    "vg_instnorm_workspace_bytes": (_Z, [_I, _I, _I, _I, _I]),
    "vg_instnorm_stats": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _Z, _P]),
    "vg_instnorm_apply": (_I, [_ID, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "vg_instnorm_bwd": (_I, [_ID, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _Z, _P]),
    "vg_instnorm_bwd_sinks": (_I, [_ID, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _Z, _P]),
    "vg_upsample_concat": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "vg_upsample_concat_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "vg_gather_pad": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "vg_gather_pad_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "vg_maxpool2_pad": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "vg_maxpool2_pad_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "vg_batchnorm_workspace_bytes": (_Z, [_I, _I, _I, _I, _I]),
    "vg_batchnorm_stats": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _F, _I, _P, _Z, _P]),
    "vg_batchnorm_bwd": (_I, [_ID, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _Z, _P]),
    "vg_conv3d_transpose_k2s2_scatter": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "vg_conv3d_transpose_k2s2_gather": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "vg_upsample_pad": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "vg_upsample_pad_bwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "vg_conv3d_transpose_k2s2_weights": (_I, [_P, _P, _I, _I, _I, _P]),
    "vg_comm_available": (_I, []),
    "vg_comm_nccl_version": (_I, []),
    "vg_comm_unique_id": (_I, [_P]),
    "vg_comm_init": (_I, [C.POINTER(C.c_void_p), _P, _I, _I, _I]),
    "vg_comm_world": (_I, [_P]),
    "vg_comm_rank": (_I, [_P]),
    "vg_comm_collectives": (_ULL, [_P]),
    "vg_comm_allreduce_bucket": (_I, [_P, _P, _LL, _LL, _P]),
    "vg_comm_reduce_scalars": (_I, [_P, _P, _I, _P]),
    "vg_comm_destroy": (_I, [_P]),
    "vg_pad_noise": (_I, [_P, _P, _I, _I, _I, _I, _P, _F, _ULL, _P, _P]),
    "vg_dropout_mask": (_I, [_P, _I, _F, _ULL, _P, _P]),
    "vg_pad_fold": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "vg_accumulate": (_I, [_P, _P, _Z, _I, _P]),
    "vg_tanh_bwd": (_I, [_P, _P, _P, _Z, _P]),
    "vg_soft_skel_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "vg_soft_skel_bwd_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "vg_soft_skel_bwd": (_I, [_P, _P, _P, _P, _P, _Z, _I, _I, _I, _I, _I, _P]),
    "vg_minmax": (_I, [_P, _I, _Z, _P, _P, _P]),
    "vg_minmax_normalize": (_I, [_P, _P, _P, _I, _Z, _P]),
    "vg_minmax_normalize_bwd": (_I, [_P, _P, _P, _P, _P, _I, _Z, _P, _I, _P]),
    "vg_sqdiff_sum": (_I, [_P, _P, _F, _Z, _P, _P]),
    "vg_lincomb": (_I, [_P, _Z, _I, _F, _P, _F, _P, _F, _P, _F, _P]),
    "vg_lincomb_dev": (_I, [_P, _Z, _I, _P, _P, _P, _P, _P]),
    "vg_cldice_coeffs": (_I, [_P, _F, _F, _P, _P]),
    "vg_bce_sum": (_I, [_P, _P, _Z, _P, _P]),
    "vg_bce_bwd": (_I, [_P, _P, _F, _P, _Z, _I, _P]),
    "vg_cldice_sums": (_I, [_P, _P, _P, _P, _Z, _P, _P]),
    "vg_ssim_fwd": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "vg_ssim_bwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P, _I, _P]),
    "vg_clip_adam_step": (_I, [_P, _P, _P, _P, _P, _I, _LL, _F, _F, _F, _F, _F, _P, _P]),
    "vg_clip_adam_step_dev": (_I, [_P, _P, _P, _P, _P, _I, _LL, _P, _F, _F, _F, _F, _P, _P]),
    "vg_stitch_gather": (_I, [_P, _I, _I, _I, _P, _P, _I, _I, _I, _I, _P]),
    "vg_crop_augment": (_I, [_P, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "vg_stitch_gather_sym": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _P]),
    "vg_stitch_accumulate": (_I, [_P, _P, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "vg_stitch_finalize": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "vg_stitch_scale": (_I, [_P, _Z, _P, _P]),
    "vg_stitch_gather_sum": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P]),
    "vg_stitch_minmax_decode": (_I, [_P, _P, _P]),
    "vg_stitch_scale_u8": (_I, [_P, _P, _Z, _P, _P]),
}

_lib = None
LAUNCHES = 0  # number of ABI compute calls issued


class Profiler:
    """Optional per-family device timing (CUDA events on the launching stream) used by bench.py to
    report the dominant kernel's achieved FLOP/s / bytes/s.  `track`: set of ABI names."""

    def __init__(self, track, detail=False):
        self.track = set(track)
        self.records = {}
        self.detail = detail   # key records by (name, descriptor shape) instead of name only

    def key(self, name, args):
        if self.detail and args:
            a = args[0]
            if isinstance(a, ConvDesc):
                return "%s %d->%d k%ds%d in%d N%d" % (name, a.Cin, a.Cout, a.K, a.stride, a.ID, a.N)
            if isinstance(a, InDesc):
                return "%s C%d S%d N%d" % (name, a.C, a.D, a.N)
        return name

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, recs in self.records.items():
            ms = sum(a.elapsed_time(b) for a, b, _ in recs)
            out[name] = dict(calls=len(recs), ms=ms, work=sum(w for _, _, w in recs))
        return out


PROFILER = None


def lib():
    """Loads the shared library (once).  Raises if it is missing — there is no CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VgError("libvangan_b200.so not built: run `python van-gan_b200/build.py` (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, int):
        return t
    assert t.is_cuda and t.is_contiguous(), "ABI buffers must be contiguous CUDA tensors"
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args, work=0.0):
    """Invoke an int-returning ABI function; tensors are converted to raw device pointers and the
    current torch stream is appended as the last argument.  `work`: algorithmic FLOPs (or bytes) of
    this call, only used when a Profiler is attached."""
    global LAUNCHES
    conv = [(_ptr(a) if (a is None or torch.is_tensor(a)) else a) for a in args]
    prof = PROFILER
    if prof is not None and (name in prof.track or prof.detail):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib(), name)(*conv, _stream())
        e1.record()
        prof.records.setdefault(prof.key(name, args), []).append((e0, e1, work))
    else:
        rc = getattr(lib(), name)(*conv, _stream())
    LAUNCHES += 1
    if rc != 0:
        raise VgError("%s failed: %s (%d)" % (name, _ERR.get(rc, "?"), rc))


def dtype_code(t):
    if t.dtype == torch.bfloat16:
        return VG_BF16
    if t.dtype == torch.float32:
        return VG_F32
    raise VgError("unsupported dtype %s" % t.dtype)
