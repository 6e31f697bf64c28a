"""ctypes binding of ``librechub_b200.so`` (C ABI declared in ``include/rechub_b200.h``).

PyTorch is used for device memory and streams only: every call passes raw ``data_ptr()`` values and the
current stream handle.  Nothing here computes.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
PACKAGE_ROOT = os.path.abspath(os.path.join(_HERE, "..", ".."))  # .../torch-rechub_b200
LIB_PATH = os.environ.get("RECHUB_B200_LIB", os.path.join(PACKAGE_ROOT, "lib", "librechub_b200.so"))
RH_MAX_FIELDS = 64
RH_MAX_DENSE = 32

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_i64 = ctypes.c_int64
c_f = ctypes.c_float


class RhField(ctypes.Structure):
    _fields_ = [("table", c_p), ("table_grad", c_p), ("ids", c_p), ("id_stride", c_i64), ("ids_are_i32", ctypes.c_int32), ("vocab", ctypes.c_int32), ("padding_idx", ctypes.c_int32), ("tile_col", ctypes.c_int32),
                ("fm_slot", ctypes.c_int32)]


class RhDense(ctypes.Structure):
    _fields_ = [("values", c_p), ("stride", c_i64), ("dtype", ctypes.c_int32), ("width", ctypes.c_int32), ("tile_col", ctypes.c_int32)]


class RhSync(ctypes.Structure):  # include/rechub_b200.h: rh_sync
    _fields_ = [("wait_flags", c_p), ("step", c_p), ("sig_flags", c_p), ("ticket", c_p), ("wait_mask", ctypes.c_int32), ("sig_world", ctypes.c_int32), ("sig_rank", ctypes.c_int32),
                ("reserved", ctypes.c_int32), ("id_snapshot_delta", c_i64)]


# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/rechub_b200.h one to one
PROTOTYPES = {
    "rh_abi_version": [],
    "rh_last_error": [],
    "rh_launch_count": [],
    "rh_l2_fetch_granularity": [c_i],
    "rh_set_pdl": [c_i],
    "rh_set_smem_carveout": [c_i],
    "rh_fields_fwd": [c_p, c_i, c_i, c_p, c_i, c_i, c_p, c_i64, c_p, c_p, c_p, c_p, c_p, c_p, c_p],
    "rh_fields_fwd_p2p": [c_p, c_i, c_i, c_i, c_p, c_i, c_i, c_i64, c_p, c_p],
    "rh_ids_scatter": [c_p, c_i, c_p, c_i, c_p, c_i, c_i, c_i64, c_i64, c_p],
    "rh_ids_scatter_signal": [c_p, c_i, c_p, c_i, c_p, c_i, c_i, c_i64, c_i64, c_p, c_i, c_i, c_p, c_p, c_p],
    "rh_fields_fwd_sync": [c_p, c_i, c_i, c_p, c_i, c_i, c_p, c_i64, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_p, c_p, c_p],
    "rh_fields_bwd": [c_p, c_i, c_i, c_i, c_p, c_i64, c_p, c_i64, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p],
    "rh_rows_gather": [c_p, c_i, c_i, c_p, c_i, c_i64, c_p, c_p, c_p],
    "rh_rows_scatter_add": [c_p, c_i, c_i, c_i, c_p, c_i, c_i64, c_p, c_p, c_p],
    "rh_seq_pool_fwd": [c_p, c_i, c_i, c_p, c_i, c_i, c_i, c_i, c_i64, c_p, c_i64, c_p, c_p],
    "rh_seq_pool_bwd": [c_p, c_i, c_i, c_i, c_p, c_i, c_i, c_i, c_i, c_i64, c_p, c_i64, c_p, c_p],
    "rh_rows_zero": [c_p, c_i, c_i, c_p, c_i, c_i64, c_p],
    "rh_rowwise_update": [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_p, c_i, c_i64, c_i, c_p, c_p, c_p, c_i64, c_f, c_f, c_f, c_f, c_p],
    "rh_opt_advance": [c_p, c_p, c_f, c_f, c_p],
    "rh_fields_rowwise_update": [c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_p, c_i64, c_f, c_f, c_f, c_f, c_p],
    "rh_fields_prefetch": [c_p, c_i, c_i, c_i, c_p, c_i64, c_p],
    "rh_fields_zero": [c_p, c_i, c_i, c_i, c_p],
    "rh_fm_fwd": [c_p, c_i, c_i, c_i, c_i, c_p, c_p],
    "rh_fm_bwd": [c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p],
    "rh_tile_fm_lr_fwd": [c_p, c_i64, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p],
    "rh_tile_fm_lr_bwd": [c_p, c_i64, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_i64, c_i, c_p, c_p, c_p],
    "rh_cross_fwd": [c_p, c_i64, c_i, c_i, c_i, c_p, c_p, c_p, c_i64, c_p, c_p],
    "rh_cross_bwd": [c_p, c_i64, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_i64, c_p, c_i64, c_p, c_p, c_p],
    "rh_colstats": [c_p, c_i64, c_i64, c_i, c_p, c_p, c_p, c_p, c_p, c_f, c_p],
    "rh_bn_act_fwd": [c_p, c_i64, c_i64, c_i, c_p, c_p, c_f, c_p, c_p, c_i, c_p, c_f, c_f, ctypes.c_uint32, c_p, c_p, c_i64, c_p],
    "rh_bn_act_bwd": [c_p, c_i64, c_i64, c_i, c_p, c_p, c_f, c_p, c_p, c_i, c_p, c_f, c_f, ctypes.c_uint32, c_p, c_p, c_i64, c_i, c_p, c_i64, c_p, c_p, c_p, c_p],
    "rh_bn_fused_scratch_floats": [c_i],
    "rh_bn_fused_supported": [c_i64, c_i, c_i],
    "rh_bn_act_fused_fwd": [c_p, c_i64, c_i64, c_i, c_f, c_p, c_p, c_i, c_p, c_f, c_f, ctypes.c_uint32, c_p, c_p, c_p, c_f, c_p, c_p, c_p, c_i64, c_p, c_p, c_p, c_p, c_i, c_p, c_p],
    "rh_bn_act_fused_bwd": [c_p, c_i64, c_i64, c_i, c_p, c_f, c_p, c_p, c_i, c_p, c_f, c_f, ctypes.c_uint32, c_p, c_i64, c_p, c_p, c_p, c_i, c_p, c_p, c_i64, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p],
    "rh_head_fwd": [c_p, c_i64, c_i64, c_i, c_p, c_p, c_p, c_p, c_i, c_p, c_p],
    "rh_head_bwd": [c_p, c_i64, c_i64, c_i, c_p, c_p, c_p, c_i, c_p, c_i64, c_p, c_p, c_p, c_p],
    "rh_gemm_tile_n": [c_i],
    "rh_gemm_options": [c_i, c_i],
    "rh_peer_wait": [c_p, c_i, c_p, c_p],
    "rh_copy_segments": [c_i, c_p, c_p, c_p, c_p],
    "rh_bce_fwd": [c_p, c_p, c_i64, c_p, c_p, c_p],
    "rh_bce_bwd": [c_p, c_p, c_p, c_i64, c_p, c_p],
    "rh_dense_update": [c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_f, c_f, c_f, c_f, c_p],
    "rh_dense_stage_floats": [c_i, c_p],
    "rh_peer_barrier": [c_p, c_p, c_i, c_i, c_p, c_p],
    "rh_dense_pack_signal": [c_i, c_p, c_p, c_p, c_i, c_p, c_p, c_i, c_i, c_p, c_p, c_p],
    "rh_dense_reduce_update": [c_i, c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_p, c_i, c_i, c_p, c_p, c_i, c_p, c_p, c_f, c_f, c_f, c_f, c_p],
    "rh_gemm_tf32x3": [c_p, c_i64, c_i, c_p, c_i64, c_i, c_p, c_i64, c_i, c_i, c_i, c_p, c_i, c_p],
    "rh_gemm_stats_scratch_floats": [c_i, c_i],
    "rh_gemm_tf32x3_stats": [c_p, c_i64, c_i, c_p, c_i64, c_i, c_p, c_i64, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_f, c_p],
    "rh_crossmix_pack": [c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_i64, c_p, c_p, c_p],
    "rh_crossmix_unpack_grads": [c_i, c_i, c_i, c_i, c_p, c_i64, c_p, c_i64, c_p, c_i64, c_p, c_p, c_p, c_p, c_p],
    "rh_crossmix_mid1_fwd": [c_p, c_i64, c_i64, c_i, c_i, c_p, c_p, c_p],
    "rh_crossmix_mid2_fwd": [c_p, c_p, c_i64, c_i, c_i, c_p, c_p, c_p],
    "rh_crossmix_out_fwd": [c_p, c_i64, c_p, c_i64, c_p, c_i64, c_p, c_i, c_i64, c_i, c_p, c_i64, c_p],
    "rh_crossmix_out_bwd": [c_p, c_i64, c_p, c_i64, c_p, c_i64, c_p, c_i64, c_p, c_i, c_i64, c_i, c_p, c_i64, c_p, c_i64, c_p, c_i64, c_p, c_p],
    "rh_crossmix_mid2_bwd": [c_p, c_i64, c_p, c_p, c_i64, c_i, c_i, c_p, c_p, c_i64, c_p],
    "rh_crossmix_mid1_bwd": [c_p, c_i64, c_p, c_i64, c_i, c_i, c_p, c_i64, c_p],
    "rh_sum3": [c_p, c_i64, c_p, c_i64, c_p, c_i64, c_i64, c_i, c_p, c_i64, c_p],
    "rh_inbatch_sample_random": [c_i, c_i, c_p, c_p, c_p],
    "rh_inbatch_sample_hard": [c_p, c_i64, c_i, c_i, c_p, c_p],
    "rh_inbatch_ce_fwd": [c_p, c_i64, c_p, c_i64, c_i, c_p, c_i, c_i, c_p, c_p, c_p],
    "rh_inbatch_ce_bwd": [c_p, c_i64, c_p, c_i64, c_i, c_p, c_p, c_p, c_i, c_i, c_p, c_i64, c_p, c_i64, c_p],
    "rh_din_attn_input_fwd": [c_p, c_i, c_p, c_i, c_i, c_p, c_p, c_i, c_i64, c_i, c_i, c_p, c_p, c_p, c_p, c_p],
    "rh_din_weighted_sum_fwd": [c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p],
    "rh_din_weighted_sum_bwd": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p],
    "rh_din_attn_input_bwd": [c_p, c_i, c_i, c_p, c_i, c_i, c_i, c_p, c_p, c_i, c_i64, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p],
}
_RESTYPES = {"rh_last_error": ctypes.c_char_p, "rh_launch_count": ctypes.c_ulonglong, "rh_gemm_stats_scratch_floats": ctypes.c_int64, "rh_bn_fused_scratch_floats": ctypes.c_int64, "rh_dense_stage_floats": ctypes.c_int64}

_lock = threading.Lock()
_lib = None


class EngineMissing(RuntimeError):
    pass


def lib():
    """The loaded library; raises :class:`EngineMissing` (loudly) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise EngineMissing("librechub_b200.so not found at %s — CUDA tensors need the sm_100a engine; build it with "
                                    "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C torch-rechub_b200/csrc`. "
                                    "There is no eager fallback for CUDA inputs." % LIB_PATH)
            handle = ctypes.CDLL(LIB_PATH)
            for name, argtypes in PROTOTYPES.items():
                fn = getattr(handle, name)  # AttributeError if the .so does not export a declared symbol
                fn.argtypes = argtypes
                fn.restype = _RESTYPES.get(name, ctypes.c_int)
            if handle.rh_abi_version() != 1:
                raise EngineMissing("librechub_b200.so ABI version %d != 1" % handle.rh_abi_version())
            from . import config
            handle.rh_set_pdl(int(config.pdl))
            handle.rh_gemm_tile_n(int(config.gemm_tile_n))
            handle.rh_set_smem_carveout(int(config.smem_carveout))
            _lib = handle
    return _lib


class EngineError(RuntimeError):
    pass


def check(status, what):
    if status != 0:
        msg = lib().rh_last_error()
        msg = msg.decode() if msg else ""
        if status == 2:
            raise NotImplementedError("%s: %s" % (what, msg))
        raise EngineError("%s failed (status %d): %s" % (what, status, msg))


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """data_ptr of a tensor, or NULL."""
    return None if t is None else t.data_ptr()


# ---- out-of-range id flag --------------------------------------------------------------------------
_err_flags = {}
l2_fetch_granularity_seen = {}  # device index -> (bytes in force before the engine touched it, bytes in force now)


def _device_init(key):
    """Once per device, at the engine's first use of it: lower the L2 fetch granularity (config.l2_fetch_granularity; 0 = leave)."""
    from . import config
    L = lib()
    with torch.cuda.device(key):
        before = int(L.rh_l2_fetch_granularity(0))
        now = before
        want = int(config.l2_fetch_granularity)
        if want > 0 and before > 0 and want != before:
            now = int(L.rh_l2_fetch_granularity(want))
    l2_fetch_granularity_seen[key] = (before, now)


def err_flag(device):
    """Per-device int32 flag the kernels set on an out-of-range id."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    t = _err_flags.get(key)
    if t is None:
        _device_init(key)
        t = torch.zeros(1, dtype=torch.int32, device=device)
        _err_flags[key] = t
    return t


def check_errors(device=None):
    """Raise the reference's ``IndexError`` if any kernel since the last check saw an out-of-range id.

    Costs one D2H read (a sync); CTRTrainer calls it where the reference already syncs (``loss.item()``).
    """
    flags = list(_err_flags.values()) if device is None else [err_flag(torch.device(device))]
    for t in flags:
        v = int(t.item())
        if v != 0:
            t.zero_()
            if v == 0x7ffffff0:  # RH_ERRFLAG_SYNC_TIMEOUT
                raise RuntimeError("a cross-GPU hand-over of the sharded exchange timed out: a peer rank never published its step (crashed or diverged rank?)")
            raise IndexError("index out of range in self (embedding lookup, field #%d of the launch)" % (v - 1))
