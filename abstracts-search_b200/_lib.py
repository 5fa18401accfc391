"""ctypes binding of libabsb200.so (include/absb200.h).

There is deliberately no fallback here: if the shared library is missing or fails to load, importing
the product path raises.  The library itself refuses to create handles on anything but an sm_100
device.
"""
from __future__ import annotations

import ctypes
import os
import re
import subprocess
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libabsb200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "absb200.h")
CSRC = os.path.join(_HERE, "csrc")

ABSB_OK = 0
ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_UNSUPPORTED, ERR_OOM = -1, -2, -3, -4, -5
MAX_K = 256


class AbsbError(RuntimeError):
    """C++ failures surface as RuntimeError, like faiss's SWIG-wrapped exceptions."""

    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


def build(verbose: bool = False, jobs: int | None = None) -> str:
    """Compile csrc/*.cu for sm_100a into libabsb200.so (nvcc cross-compiles without a GPU)."""
    jobs = jobs or min(8, os.cpu_count() or 1)
    r = subprocess.run(["make", "-C", CSRC, f"-j{jobs}"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-8000:])
    if r.returncode != 0:
        raise RuntimeError("building libabsb200.so failed")
    return LIB_PATH


def header_functions() -> list[str]:
    """Names of every function include/absb200.h declares."""
    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(absb_[a-z0-9_]+)\s*\(", src)))


class EncConfig(ctypes.Structure):
    _fields_ = [
        ("vocab_size", c_int32), ("hidden_size", c_int32), ("num_layers", c_int32),
        ("num_heads", c_int32), ("num_kv_heads", c_int32), ("head_dim", c_int32),
        ("intermediate_size", c_int32), ("embed_dim", c_int32), ("max_seq_len", c_int32),
        ("causal", c_int32), ("rms_eps", c_float), ("rope_theta", c_float),
    ]


_PF, _PI64, _PI32, _PD = POINTER(c_float), POINTER(c_int64), POINTER(c_int32), POINTER(c_double)
_H = c_void_p  # opaque handles

_SIGS = {
    "absb_version": ([], c_int),
    "absb_last_error": ([], c_char_p),
    "absb_device_count": ([POINTER(c_int)], c_int),
    "absb_device_info": ([c_int, c_char_p, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int)], c_int),
    "absb_synth_fill_dev": ([c_int, c_uint64, c_int64, c_int64, c_int, c_int, c_int64, c_void_p, c_void_p], c_int),
    "absb_synth_fill_rows_dev": ([c_int, c_uint64, c_void_p, c_int64, c_int, c_int, c_int64, c_void_p, c_void_p], c_int),
    "absb_synth_cluster_dev": ([c_uint64, c_int64, c_int64, c_int, c_void_p, c_void_p], c_int),
    # flat
    "absb_flat_create": ([c_int, c_int, c_int, POINTER(_H)], c_int),
    "absb_flat_destroy": ([_H], c_int),
    "absb_flat_reset": ([_H], c_int),
    "absb_flat_ntotal": ([_H, _PI64], c_int),
    "absb_flat_add": ([_H, c_int64, c_void_p], c_int),
    "absb_flat_add_dev": ([_H, c_int64, c_void_p, c_void_p], c_int),
    "absb_flat_search": ([_H, c_int64, c_void_p, c_int, c_void_p, c_void_p], c_int),
    "absb_flat_search_dev": ([_H, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_void_p], c_int),
    "absb_flat_reconstruct": ([_H, c_int64, c_int64, c_void_p], c_int),
    # ivf
    "absb_ivf_create": ([c_int, c_int, c_int, c_int, POINTER(_H)], c_int),
    "absb_ivf_destroy": ([_H], c_int),
    "absb_ivf_reset": ([_H], c_int),
    "absb_ivf_ntotal": ([_H, _PI64], c_int),
    "absb_ivf_is_trained": ([_H, POINTER(c_int)], c_int),
    "absb_ivf_set_clustering": ([_H, c_int, c_int, c_int, c_int64], c_int),
    "absb_ivf_set_clustering_spherical": ([_H, c_int], c_int),
    "absb_renorm_rows_dev": ([c_int, c_int64, c_int, c_void_p, c_void_p], c_int),
    "absb_ivf_train": ([_H, c_int64, c_void_p], c_int),
    "absb_ivf_train_dev": ([_H, c_int64, c_void_p, c_void_p], c_int),
    "absb_ivf_set_centroids": ([_H, c_void_p], c_int),
    "absb_ivf_get_centroids": ([_H, c_void_p], c_int),
    "absb_ivf_set_centroids_dev": ([_H, c_void_p, c_void_p], c_int),
    "absb_ivf_add": ([_H, c_int64, c_void_p, c_void_p], c_int),
    "absb_ivf_add_dev": ([_H, c_int64, c_void_p, c_void_p, c_void_p], c_int),
    "absb_ivf_add_preassigned": ([_H, c_int64, c_void_p, c_void_p, c_void_p], c_int),
    "absb_ivf_add_preassigned_dev": ([_H, c_int64, c_void_p, c_void_p, c_void_p, c_void_p], c_int),
    "absb_ivf_compact": ([_H], c_int),
    "absb_ivf_compact_scratch": ([_H, c_int64], c_int),
    "absb_plan_page_compaction": ([c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, _PI64, _PI64], c_int),
    "absb_ivf_coarse": ([_H, c_int64, c_void_p, c_int, c_void_p, c_void_p], c_int),
    "absb_ivf_coarse_dev": ([_H, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_void_p], c_int),
    "absb_ivf_assign": ([_H, c_int64, c_void_p, c_void_p], c_int),
    "absb_ivf_search": ([_H, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p], c_int),
    "absb_ivf_search_dev": ([_H, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p], c_int),
    "absb_ivf_search_preassigned": ([_H, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p], c_int),
    "absb_ivf_search_preassigned_dev": ([_H, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p], c_int),
    "absb_ivf_list_sizes": ([_H, c_void_p], c_int),
    "absb_ivf_get_list": ([_H, c_int64, c_void_p, c_void_p], c_int),
    "absb_ivf_centroid_sums_dev": ([_H, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p], c_int),
    "absb_rand_perm": ([c_int64, c_int64, c_void_p], c_int),
    "absb_kmeans_split_clusters": ([c_int, c_int64, c_int64, c_void_p, c_void_p, _PI64], c_int),
    "absb_ivf_set_shard": ([_H, c_int, c_int], c_int),
    "absb_merge_shards_dev": ([c_int, c_int, c_int64, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p], c_int),
    "absb_ivf_set_tunables": ([_H, c_int, c_int, c_int], c_int),
    "absb_ivf_set_scan_order": ([_H, c_int], c_int),
    "absb_ivf_set_scan_impl": ([_H, c_int, c_int, c_int, c_int], c_int),
    "absb_ivf_set_two_stage": ([_H, c_int], c_int),
    "absb_ivf_two_stage_fallbacks": ([_H, _PI64], c_int),
    "absb_ivf_last_stats": ([_H, _PI64, _PI64, _PI64, _PI64], c_int),
    "absb_ivf_set_profile": ([_H, c_int], c_int),
    "absb_ivf_get_profile": ([_H, _PD, _PD, _PD, _PI64], c_int),
    "absb_ivf_time_scan": ([_H, c_int, c_void_p, _PF], c_int),
    "absb_ivf_get_profile_scan16": ([_H, _PD, _PI64], c_int),
    "absb_ivf_profile_spans": ([_H, c_void_p, c_void_p, c_int64, _PI64], c_int),
    "absb_enc_profile_spans": ([_H, c_void_p, c_void_p, c_int64, _PI64], c_int),
    # encoder
    "absb_enc_create": ([POINTER(EncConfig), c_int, POINTER(_H)], c_int),
    "absb_enc_destroy": ([_H], c_int),
    "absb_enc_load_weight": ([_H, c_char_p, c_void_p, c_int, _PI64, c_int], c_int),
    "absb_enc_init_random": ([_H, c_uint64, c_float], c_int),
    "absb_enc_get_weight": ([_H, c_char_p, c_void_p, c_int64], c_int),
    "absb_enc_forward": ([_H, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p], c_int),
    "absb_enc_forward_dev": ([_H, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p], c_int),
    "absb_enc_last_hidden": ([_H, c_void_p, c_int64], c_int),
    "absb_enc_last_stats": ([_H, _PD, _PI64], c_int),
    "absb_enc_set_attention_impl": ([_H, c_int], c_int),
    "absb_enc_set_profile": ([_H, c_int], c_int),
    "absb_enc_get_profile": ([_H, _PD, _PD, _PD, _PD, _PI64], c_int),
    "absb_gemm_set_variant": ([c_int], c_int),
    "absb_gemm_set_ksplit": ([c_int], c_int),
    "absb_gemm_set_smem_budget": ([c_int], c_int),
    "absb_gemm_bf16_epi_dev": ([c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p], c_int),
    "absb_gemm_bf16_dev": ([c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p], c_int),
    # NVLink peer exchange
    "absb_peer_create": ([c_int, c_int, c_int, c_size_t, POINTER(_H)], c_int),
    "absb_peer_destroy": ([_H], c_int),
    "absb_peer_ipc_handle": ([_H, c_void_p], c_int),
    "absb_peer_local_ptr": ([_H, POINTER(c_void_p)], c_int),
    "absb_peer_connect": ([_H, c_void_p], c_int),
    "absb_peer_connect_ptrs": ([_H, c_void_p], c_int),
    "absb_peer_allgather_dev": ([_H, c_void_p, c_size_t, POINTER(c_void_p), c_void_p], c_int),
    "absb_peer_push_dev": ([_H, c_void_p, c_size_t, c_void_p], c_int),
    "absb_peer_wait_dev": ([_H, POINTER(c_void_p), c_void_p], c_int),
    "absb_peer_status": ([_H, POINTER(c_int)], c_int),
    "absb_ivf_search_push_dev": ([_H, _H, c_int64, c_void_p, c_int, c_int, c_void_p], c_int),
    "absb_ivf_search_preassigned_push_dev": ([_H, _H, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p], c_int),
    "absb_peer_push_results_dev": ([_H, c_int64, c_int, c_void_p, c_void_p, c_void_p], c_int),
    "absb_peer_merge_shards_dev": ([_H, c_int64, c_int, c_void_p, c_void_p, c_void_p], c_int),
    # OpenAlex JSON-lines front end (host only)
    "absb_oa_jsonl_convert": ([c_char_p, c_size_t, c_int, c_int, POINTER(c_char_p), POINTER(c_size_t),
                               POINTER(c_size_t), _PI64], c_int),
    "absb_oa_jsonl_free": ([c_char_p], c_int),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load libabsb200.so once.  Raises if it is not built — the product path has no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C abstracts-search_b200/csrc`). There is no CPU/PyTorch fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGS.items():
            fn = getattr(handle, name)  # AttributeError if the library lacks a declared symbol
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc == ABSB_OK:
        return
    msg = lib().absb_last_error().decode("utf-8", "replace")
    if rc == ERR_OOM:
        raise MemoryError(msg)
    raise AbsbError(rc, msg)


def ptr(a) -> c_void_p:
    """Raw pointer of a numpy array (host) or a torch tensor (device or host)."""
    if a is None:
        return c_void_p(0)
    if hasattr(a, "data_ptr"):
        return c_void_p(a.data_ptr())
    return c_void_p(a.ctypes.data)


def current_stream_ptr() -> c_void_p:
    import torch

    return c_void_p(torch.cuda.current_stream().cuda_stream)
