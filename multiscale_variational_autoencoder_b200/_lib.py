"""ctypes binding of libmvae_b200.so (the C-ABI declared in include/mvae_b200.h).

There is no CPU fallback: importing the package without the built library, or calling an op on a device that
is not sm_100, raises.  `build()` compiles the library in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libmvae_b200.so")
SOURCES = ["api.cu", "pyramid.cu", "pyramid_tiled.cu", "pyramid_small.cu", "elbo.cu", "conv_simt.cu", "conv_small_cin.cu", "conv_tc.cu", "dense_tc.cu", "blocks.cu", "dwconv_tiled.cu", "se_gate.cu", "mbv3_fused.cu", "optim.cu", "comm.cu"]

ACT_NONE, ACT_RELU, ACT_ELU = 0, 1, 2
DIFF_NO_UPSAMPLE, DIFF_LAPLACIAN = 0, 1
PREC_FP32, PREC_TF32 = 0, 1
REG_NONE, REG_L1, REG_L2 = 0, 1, 2


class MvaeError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("B", "H", "W", "Cin", "kh", "kw", "sh", "sw", "Cout", "coord_mode", "precision")]


class Mbv3FwdArgs(C.Structure):
    """mvae_mbv3_fwd_args (include/mvae_b200.h)"""
    _fields_ = [(n, C.c_int) for n in ("B", "H", "W", "C")] + [(n, C.c_void_p) for n in (
        "u_prev", "x_prev", "gate_prev", "w2", "b2", "y", "x", "w0", "b0", "wd", "bd", "a", "u", "gap_sum",
        "se_w0", "se_b0", "se_ws", "se_stat", "se_stat_prev", "se_gamma_prev", "se_beta_prev", "se_w1_prev", "se_b1_prev", "se_mm_prev", "se_mv_prev",
        "se_ws_prev", "gate_out_prev")] + [("bn_eps", C.c_float), ("bn_momentum", C.c_float), ("training", C.c_int)]


class Mbv3BwdArgs(C.Structure):
    """mvae_mbv3_bwd_args (include/mvae_b200.h)"""
    _fields_ = [(n, C.c_int) for n in ("B", "H", "W", "C")] + [(n, C.c_void_p) for n in (
        "dy", "u", "a", "gate", "dgap", "w2", "wd", "w0", "da", "dx", "dwd", "dbd", "w2_prev", "u_prev", "dgate_prev",
        "se_w1_prev", "se_ws_prev", "se_bstat_prev", "se_w0", "se_gamma", "se_ws", "se_bstat")]


def _source_hash():
    import hashlib
    h = hashlib.sha256()
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(_HERE, "..", "include", "mvae_b200.h")]
    for d in deps:
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _needs_build():
    """Stale when the sources' content hash differs from the one recorded at build time (mtimes do not survive the
    copy to the GPU box)."""
    if not os.path.exists(LIB_PATH) or not os.path.exists(LIB_PATH + ".hash"):
        return True
    with open(LIB_PATH + ".hash") as f:
        return f.read().strip() != _source_hash()


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a -> libmvae_b200.so (in-tree, travels with gpurun)."""
    if not force and not _needs_build():
        return LIB_PATH
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise MvaeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    with open(LIB_PATH + ".hash", "w") as f:
        f.write(_source_hash())
    return LIB_PATH


_P, _I, _F, _LL, _SZ = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_size_t
_PD = C.POINTER(ConvDesc)

# name -> (restype, argtypes); must list every symbol of include/mvae_b200.h (checked by tests/test_abi.py)
PROTOTYPES = {
    "mvae_version": (_I, []),
    "mvae_last_error": (_I, [C.c_char_p, _SZ]),
    "mvae_device_arch": (_I, []),
    "mvae_memset_zero": (_I, [_P, _SZ, _P]),
    "mvae_accumulate": (_I, [_P, _P, _I, _F, _P]),
    "mvae_tc_launch_count": (_LL, []),
    "mvae_kernel_launch_count": (_LL, []),
    "mvae_set_wgrad_sm_share": (_I, [_I]),
    "mvae_stream_create": (_I, [_I, C.POINTER(_P)]),
    "mvae_stream_destroy": (_I, [_P]),
    "mvae_debug_trace": (_I, [_P]),
    "mvae_pyramid_split_workspace_bytes": (_SZ, [_I] * 5),
    "mvae_pyramid_split": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _F, _F, _P, _I, _I, _I, _P]),
    "mvae_gaussian_filter": (_I, [_P, _P, _I, _I, _I, _I, _P, _I, _I, _P]),
    "mvae_input_corrupt": (_I, [_P, _P, _P, _P, _I, _I, _I, _F, _F, _F, _F, _P]),
    "mvae_pyramid_merge_workspace_bytes": (_SZ, [_I] * 5),
    "mvae_pyramid_merge_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "mvae_pyramid_merge_bwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "mvae_denormalize_clip": (_I, [_P, _P, _LL, _F, _F, _P]),
    "mvae_recon_loss_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _P]),
    "mvae_recon_loss_bwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _F, _P]),
    "mvae_loss_finalize": (_I, [_P, _P, _I, _P, _P, _I, _I, _I, _I, _F, _F, _P]),
    "mvae_reparam_kl_fwd": (_I, [_P, _P, _P, _P, _I, _I, _F, _F, _P]),
    "mvae_reparam_kl_bwd": (_I, [_P, _P, _P, _P, _I, _I, _F, _F, _F, _P]),
    "mvae_conv2d_fwd": (_I, [_PD, _P, _P, _P, _P, _P, _I, _P, _P]),
    "mvae_conv2d_dgrad": (_I, [_PD, _P, _P, _P, _P, _P, _I, _P, _P]),
    "mvae_conv2d_wgrad": (_I, [_PD, _P, _P, _P, _P, _P, _P]),
    "mvae_dense_workspace_bytes": (_SZ, [_I] * 3),
    "mvae_dense_fwd": (_I, [_I, _I, _I, _P, _P, _P, _I, _P, _P, _SZ, _I, _P]),
    "mvae_dense_dgrad": (_I, [_I, _I, _I, _P, _P, _P, _I, _P, _P, _SZ, _I, _P]),
    "mvae_conv2d_fwd_batched": (_I, [_I, _P, _P, _P, _P, _P, _P, _I, _P, _P]),
    "mvae_conv2d_dgrad_batched": (_I, [_I, _P, _P, _P, _P, _P, _P, _I, _P, _P]),
    "mvae_conv2d_wgrad_batched": (_I, [_I, _P, _P, _P, _P, _P, _P, _P]),
    "mvae_dwconv3x3_fwd_batched": (_I, [_I, _P, _P, _P, _P, _P, _I, _P, _P, _I, _P]),
    "mvae_dwconv3x3_bwd_batched": (_I, [_I] + [_P] * 9 + [_I, _P, _P, _I, _P]),
    "mvae_se_gate_fwd_batched": (_I, [_I] + [_P] * 11 + [_I, _I, _P, _F, _F, _I, _P]),
    "mvae_se_gate_bwd_batched": (_I, [_I] + [_P] * 13 + [_I, _I, _P, _P]),
    "mvae_se_dgate_reduce_batched": (_I, [_I, _P, _P, _P, _I, _P, _I, _P]),
    "mvae_dwconv3x3_fwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "mvae_dwconv3x3_bwd": (_I, [_P] * 9 + [_I, _I, _I, _I, _P]),
    "mvae_se_gate_ws_floats": (_LL, [_I, _I]),
    "mvae_se_gate_fwd": (_I, [_P] * 11 + [_I, _I, _I, _F, _F, _I, _P]),
    "mvae_se_dgate_reduce": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "mvae_se_gate_bwd": (_I, [_P] * 13 + [_I, _I, _I, _P]),
    "mvae_mbv3_fused_supported": (_I, [_I] * 5),
    "mvae_mbv3_fused_fwd": (_I, [C.POINTER(Mbv3FwdArgs), _P]),
    "mvae_mbv3_fused_bwd": (_I, [C.POINTER(Mbv3BwdArgs), _P]),
    "mvae_mbv3_fused_fwd_batched": (_I, [_I, C.POINTER(Mbv3FwdArgs), _P]),
    "mvae_mbv3_fused_bwd_batched": (_I, [_I, C.POINTER(Mbv3BwdArgs), _P]),
    "mvae_channel_scale": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "mvae_colsum": (_I, [_P, _P, _LL, _I, _P]),
    "mvae_bn_stats": (_I, [_P, _P, _LL, _I, _P]),
    "mvae_bn_convout_fwd": (_I, [_P] * 10 + [_LL, _I, _I, _F, _F, _I, _P]),
    "mvae_bn_convout_bwd": (_I, [_P] * 12 + [_LL, _I, _I, _P]),
    "mvae_optim_norms": (_I, [_P, _P, _P, _P, _I, _I, _I, _F, _P, _P, _P, _P]),
    "mvae_optim_adagrad": (_I, [_P, _P, _P, _P, _P, _I, _I, _P, _P, _F, _F, _P]),
    "mvae_coord_channels": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "mvae_comm_handle_bytes": (_SZ, []),
    "mvae_comm_alloc_signals": (_I, [C.POINTER(_P)]),
    "mvae_comm_free_signals": (_I, [_P]),
    "mvae_comm_export": (_I, [_P, _P, C.POINTER(C.c_ulonglong)]),
    "mvae_comm_open": (_I, [_P, C.c_ulonglong, C.POINTER(_P), C.POINTER(_P)]),
    "mvae_comm_close": (_I, [_P]),
    "mvae_comm_allreduce": (_I, [_P, _P, _I, _I, _I, _P, _P, _I, _I, _P]),
    "mvae_comm_status": (_I, [_P, C.POINTER(_I)]),
}

_lib = None


def load():
    """Load (building first if the .so is missing or stale and nvcc is available).  Raises when impossible."""
    global _lib
    if _lib is not None:
        return _lib
    if _needs_build():
        try:
            build()
        except FileNotFoundError as e:   # no nvcc on this machine
            if not os.path.exists(LIB_PATH):
                raise MvaeError("libmvae_b200.so is not built and nvcc is unavailable; there is no CPU path") from e
            import warnings
            warnings.warn("libmvae_b200.so does not match the sources under csrc/ (content hash differs) and nvcc is "
                          "unavailable to rebuild it: loading the stale library", RuntimeWarning, stacklevel=2)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def last_error():
    buf = C.create_string_buffer(512)
    load().mvae_last_error(buf, 512)
    return buf.value.decode()


def check(rc, what=""):
    if rc != 0:
        raise MvaeError(f"{what} failed ({rc}): {last_error()}")


class OwnedStream:
    """A CUDA stream of our own (mvae_stream_create) seen by PyTorch as an ExternalStream.  torch.cuda.Stream() hands out
    streams from a pool of 32 per priority and device, round-robin: an engine that forks into ~50 streams gets aliases, and
    two branches of the step on one aliased stream serialise in the captured graph."""

    @staticmethod
    def create(device, priority=0):
        """priority: 0 = lowest (the default of CUDA streams), 1 = middle, 2 = highest (True counts as highest)."""
        import torch
        h = C.c_void_p()
        prio = 2 if priority is True else int(priority)
        with torch.cuda.device(device):
            check(load().mvae_stream_create(prio, C.byref(h)), "mvae_stream_create")
        return torch.cuda.ExternalStream(h.value, device=device)


_arch_ok = {}


def require_b200(device_index):
    """Fail loudly unless the current CUDA device is sm_100 (B200)."""
    if device_index in _arch_ok:
        return
    import torch
    if not torch.cuda.is_available():
        raise MvaeError("no CUDA device: multiscale_variational_autoencoder_b200 has no CPU path")
    with torch.cuda.device(device_index):
        arch = load().mvae_device_arch()
    if arch != 100:
        raise MvaeError(f"device sm_{arch} is not a B200 (sm_100a); the kernels are built for sm_100a only")
    _arch_ok[device_index] = True
