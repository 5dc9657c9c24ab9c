"""ctypes binding of libsgcn_b200.so (the C ABI declared in include/sgcn_b200.h).

There is no CPU fallback: if the library is missing and cannot be built, or a call fails, this
module raises.  PyTorch is used by callers only for device memory / streams; the library itself
takes raw device pointers.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsgcn_b200.so")

SGCN_OK, SGCN_EINVAL, SGCN_ECUDA, SGCN_ESTATE, SGCN_EDATA = 0, -1, -2, -3, -4

# vector ids of sgcn_sampler_vec (include/sgcn_b200.h)
VEC_FIELD, VEC_FFIELD, VEC_EDG_S, VEC_EDG_T, VEC_FEDG_S, VEC_FEDG_T = 0, 1, 2, 3, 4, 5
VEC_ADJ_I, VEC_ADJ_P, VEC_ROWPTR_S, VEC_ROWPTR_F, VEC_TGT, VEC_META = 6, 7, 10, 11, 12, 13
VEC_PIPE = 14
VEC_SCALES, VEC_EDG_W, VEC_MEDG_W, VEC_FEDG_W, VEC_ADJ_W, VEC_IMPORTANCE = 100, 101, 102, 103, 104, 105
FLOAT_VECS = {VEC_SCALES, VEC_EDG_W, VEC_MEDG_W, VEC_FEDG_W, VEC_ADJ_W, VEC_IMPORTANCE}


class SgcnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libsgcn_b200: %s (code %d)" % (msg, code))
        self.code = code


_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64

# name -> (restype, argtypes); every symbol include/sgcn_b200.h declares
SIGNATURES = {
    "sgcn_abi_version": (_i32, []),
    "sgcn_last_error": (C.c_char_p, []),
    "sgcn_launch_count": (_i64, []),
    "sgcn_trace_set": (_i32, [_vp]),
    "sgcn_sampler_create": (_i32, [C.POINTER(_vp), _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32]),
    "sgcn_sampler_create_device": (_i32, [C.POINTER(_vp), _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32]),
    "sgcn_sampler_destroy": (None, [_vp]),
    "sgcn_sampler_seed": (_i32, [_vp, _i32]),
    "sgcn_sampler_reserve": (_i32, [_vp, _i32, _vp, _i32, _i32]),
    "sgcn_sampler_start_batch": (_i32, [_vp, _i32, _vp]),
    "sgcn_sampler_start_batch_device": (_i32, [_vp, _i32, _vp]),
    "sgcn_sampler_expand": (_i32, [_vp, _i32, _i32]),
    "sgcn_sampler_sizes": (_i32, [_vp, _i32, _vp]),
    "sgcn_sampler_vec": (_i32, [_vp, _i32, _i32, C.POINTER(_vp), C.POINTER(_i64)]),
    "sgcn_sampler_copy_vec": (_i32, [_vp, _i32, _i32, _vp, _i64]),
    "sgcn_sampler_set_stream": (_i32, [_vp, _vp]),
    "sgcn_sampler_set_stream_async": (_i32, [_vp, _vp]),
    "sgcn_sampler_get_rng": (_i32, [_vp, _vp, C.POINTER(_i32)]),
    "sgcn_sampler_set_rng": (_i32, [_vp, _vp, _i32]),
    "sgcn_gather_rows": (_i32, [_vp, _i64, _vp, _i32, _vp, _i32, _vp, _i64, _vp]),
    "sgcn_csr_slice_indptr": (_i32, [_vp, _vp, _i32, _vp, _vp]),
    "sgcn_csr_slice": (_i32, [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "sgcn_spmm_csr": (_i32, [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _i64, _i32, _vp, _i64, _i32, _vp]),
    "sgcn_spmm_coo": (_i32, [_vp, _vp, _i32, _vp, _i64, _i32, _vp, _i64, _i32, _vp]),
    "sgcn_spmm_csr_bwd": (_i32, [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _i64, _i32, _vp, _i64, _vp]),
    "sgcn_full_history_mean": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _i64,
                                      _vp, _i64, _vp, _vp]),
    "sgcn_tune_set": (_i32, [_i32, _i32]),
    "sgcn_ln_act_fwd": (_i32, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, C.c_float, _i32, _vp, _i64, _vp, _vp]),
    "sgcn_ln_act_bwd": (_i32, [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _vp, _i32, _vp, _vp, _i32, _vp, _i64,
                               _vp, _vp, _vp]),
    "sgcn_dropout": (_i32, [_vp, _i64, _i32, _vp, _i32, C.c_float, C.c_uint64, C.c_uint64, _vp, _vp, _vp, _i64,
                            _vp]),
    "sgcn_xent": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _i64, _vp]),
    "sgcn_adam_step": (_i32, [_vp, _vp, _vp, _vp, _i64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                              _vp]),
    "sgcn_csr_spmm": (_i32, [_vp, _vp, _vp, _i32, _vp, _i64, _i32, _vp, _i64, _i32, _vp]),
    "sgcn_spmm_csr_sq": (_i32, [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _i64, _i32, _vp, _i64, _i32, _vp]),
    "sgcn_spmm_csr_bwd_sq": (_i32, [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _i64, _i32, _vp, _i64, _vp]),
    "sgcn_full_history_mean_sq": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _i64,
                                         _vp, _i64, _vp, _vp]),
    "sgcn_det_sampled_fwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _i64, _vp, _i64, _i32,
                                    _vp, _i64, _vp, _i64, _vp, _i64, _i32, _vp]),
    "sgcn_det_sampled_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _i64, _vp, _i64, _i32,
                                    _vp, _i64, _vp, _i64, _vp, _i64, _vp]),
    "sgcn_sampler_set_slot": (_i32, [_vp, _i32]),
    "sgcn_sampler_reserve_sets": (_i32, [_vp, _i32, _i32, _i32]),
    "sgcn_sampler_expand_train": (_i32, [_vp, _vp, _i32, _i32, _vp, _i32, _vp]),
    "sgcn_sampler_pipeline": (_i32, [_vp, _i32]),
    "sgcn_sampler_mark_consumed": (_i32, [_vp, _vp]),
    "sgcn_cv_sampled_fwd": (_i32, [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _i64, _vp, _i64, _i32, _vp,
                                   _i64, _vp, _i64, _i32, _vp]),
    "sgcn_cvd_sampled_fwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _i64, _vp, _i64, _vp,
                                    _i64, _i32, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _vp]),
    "sgcn_history_update": (_i32, [_vp, _i64, _vp, _i32, _vp, _vp, _i64, _i32, _vp, _vp]),
    "sgcn_copy_rows_pad_pair": (_i32, [_vp, _i64, _i32, _vp, _i32, _i32, _vp, _i64,
                                       _vp, _i64, _i32, _vp, _i32, _i32, _vp, _i64, _vp]),
    "sgcn_gather_pad_pair": (_i32, [_vp, _i64, _vp, _i32, _vp, _i32, _vp, _i64,
                                    _vp, _i64, _i32, _vp, _i32, _i32, _vp, _i64,
                                    _vp, _i64, _i32, _vp, _i32, _i32, _vp, _i64, _vp]),
    "sgcn_cv_sampled_fwd_bwd": (_i32, [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _i64, _vp, _i64, _i32, _vp,
                                       _i64, _vp, _i64, _i32, _vp, _i64, _vp, _i64, _vp]),
    "sgcn_cvd_sampled_fwd_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _i64, _vp, _i64, _vp,
                                        _i64, _i32, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i32,
                                        _vp, _i64, _vp, _i64, _vp]),
    "sgcn_copy_rows_pad": (_i32, [_vp, _i64, _i32, _vp, _i32, _i32, _vp, _i64, _vp]),
    "sgcn_sampler_slot_vec": (_i32, [_vp, _i32, _i32, C.POINTER(_vp)]),
    "sgcn_step_create": (_i32, [C.POINTER(_vp), _vp, _vp]),
    "sgcn_step_destroy": (None, [_vp]),
    "sgcn_step_run": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp]),
    "sgcn_step_run_trains": (_i32, [_vp, _vp, _i32, _i32, _vp, _i32, _vp]),
    "sgcn_step_status": (_i32, [_vp, C.POINTER(_i32)]),
    "sgcn_full_history_mean_wb": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _i64,
                                         _vp, _i64, _vp, _vp, _i32, _vp, _i64, _vp, _vp, _vp]),
    "sgcn_wb_counters_reset": (_i32, [_vp, _vp]),
    "sgcn_sampled_done_attach": (_i32, [_vp]),
    "sgcn_wb_payload_bytes": (_i64, [_i32, _i32]),
    "sgcn_wb_pack": (_i32, [_vp, _vp, _i32, _vp, _i64, _i32, C.POINTER(_vp), _i32, _i32, _vp]),
    "sgcn_wb_push": (_i32, [_vp, _vp, _i32, _vp, _i64, _i32, C.POINTER(_vp), C.POINTER(_vp), _i32,
                            C.POINTER(_vp), _i32, _vp, _vp, _vp]),
    "sgcn_wb_push_attach": (_i32, [_vp, _vp, _i32, _vp, _i64, _i32, C.POINTER(_vp), C.POINTER(_vp), _i32,
                                   C.POINTER(_vp), _i32, _vp, _vp]),
    "sgcn_wb_wait_apply": (_i32, [_vp, _i64, _i32, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sgcn_wb_apply": (_i32, [_vp, _i64, _i32, _vp, _i64, _i32, _i32, _vp, _vp]),
    "sgcn_shard_set": (_i32, [_i32, _i32, _i32, C.POINTER(_vp)]),
    "sgcn_wb_wait_apply_sharded": (_i32, [_vp, _i64, _i32, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _i64,
                                          _vp, _vp, _vp, C.POINTER(_vp), _vp, C.POINTER(_vp), _vp, _vp, _vp, _vp]),
    "sgcn_wb_push_ring": (_i32, [_vp, _vp, _i32, _vp, _i64, _i32, C.POINTER(_vp), _i32, _i32, _i64, C.POINTER(_vp),
                                 _i32, _vp, _vp, _vp]),
    "sgcn_wb_wait_apply_ring": (_i32, [_vp, _i64, _i32, _vp, _i64, _i32, _i32, _vp, _vp, _i32, _i64, _vp, _vp, _vp,
                                       _vp, _vp]),
    "sgcn_wb_claim_ring": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _i32, _i64, _vp, _vp, _vp, _vp]),
    "sgcn_wb_copy_ring": (_i32, [_vp, _i64, _i32, _vp, _i64, _i32, _i32, _vp, _i32, _i64, _vp, _vp, _vp, _vp]),
    "sgcn_gemm_packed_floats": (_i64, [_i32, _i32]),
    "sgcn_gemm_pack_w": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp]),
    "sgcn_gather_gemm_tf32x3": (_i32, [_vp, _i64, _vp, _i32, _vp, _i32, _vp, _i32, _vp, _i64, _vp, _i64, _vp, _i32,
                                       C.c_float, _vp]),
    "sgcn_ipc_alloc": (_i32, [C.POINTER(_vp), _i64, _i32]),
    "sgcn_ipc_free": (_i32, [_vp]),
    "sgcn_ipc_export": (_i32, [_vp, _vp]),
    "sgcn_ipc_open": (_i32, [_vp, C.POINTER(_vp)]),
    "sgcn_ipc_close": (_i32, [_vp]),
}

class StepDesc(C.Structure):
    """mirror of sgcn_step_desc (include/sgcn_b200.h)"""
    _fields_ = [("mode", _i32), ("concat", _i32), ("batch", _i32), ("degree", _i32), ("hidden", _i32),
                ("feat_dim", _i32), ("x0_rows", _i32), ("world", _i32), ("rank", _i32), ("wb_bound", _i32),
                ("features", _vp), ("ld_feat", _i64), ("history", _vp), ("ld_hist", _i64),
                ("x0", _vp), ("ld_x0", _i64), ("out", _vp * 2), ("out_mu", _vp * 2), ("ld_out", _i64),
                ("d_out", _vp), ("ld_dout", _i64), ("dx", _vp), ("ld_dx", _i64), ("slot_bytes", _i64),
                ("dst_even", _vp * 16), ("dst_odd", _vp * 16), ("peer_flags", _vp * 16),
                ("recv_even", _vp), ("recv_odd", _vp), ("flags", _vp), ("epoch", _vp), ("timeout_flag", _vp),
                ("block_counter", _vp), ("owner", _vp),
                ("x0_alt", _vp * 2), ("dx_alt", _vp), ("train", _i32), ("fuse_write_back", _i32),
                ("ring", _i32), ("pad0", _i32), ("ring_stride", _i64), ("push_epoch", _vp), ("apply_epoch", _vp),
                ("apply_stash", _vp), ("ring_flags", _vp), ("ring_dst", _vp * 16), ("ring_peer_flags", _vp * 16),
                ("ring_recv", _vp),
                ("shard_rows", _i32), ("pad1", _i32), ("hist_shards", _vp * 16), ("feat_shards", _vp * 16),
                ("reads_flags", _vp), ("reads_peer_flags", _vp * 16), ("applied_flags", _vp),
                ("applied_peer_flags", _vp * 16), ("shard_counter", _vp)]


_lib = None


def load():
    """Load (building first if the .so is absent and nvcc is available).  Raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build
        build.build_library()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    # optional overrides of the kernel tunables (include/sgcn_b200.h: SGCN_TUNE_*), for A/B runs
    for key, env in enumerate(("SGCN_FULL_VARIANT", "SGCN_TMA_WARPS", "SGCN_TMA_ROWS", "SGCN_TMA_DEPTH",
                               "SGCN_TMA_GRID", "SGCN_PDL", "SGCN_HIST_L2", "SGCN_STREAM_L2",
                               "SGCN_FULL_TRIGGER", "SGCN_FULL_REGS", "SGCN_WB_TRIGGER")):
        if os.environ.get(env, "") != "":
            if lib.sgcn_tune_set(key, int(os.environ[env])) != SGCN_OK:
                raise SgcnError(SGCN_EINVAL, "%s=%s: %s" % (env, os.environ[env], lib.sgcn_last_error().decode()))
    return lib


def last_error():
    return load().sgcn_last_error().decode("utf-8", "replace")


def check(rc):
    if rc != SGCN_OK:
        raise SgcnError(rc, last_error())


def launch_count():
    return int(load().sgcn_launch_count())


def ptr(t):
    """Raw address of a torch tensor / numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def stream_ptr(stream=None):
    """cudaStream_t of a torch stream (default: torch's current stream) as void*."""
    import torch
    if stream is None:
        stream = torch.cuda.current_stream()
    return C.c_void_p(stream.cuda_stream)
