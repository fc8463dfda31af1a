"""ctypes binding of libngsid.so (include/ngsid.h). No torch, no fallback: if the shared library
or a CUDA device is missing the import of the hot path fails loudly."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libngsid.so")

EXPORTS = ["ngsid_version", "ngsid_ctx_create", "ngsid_ctx_destroy", "ngsid_last_error",
           "ngsid_launch_count", "ngsid_poa_cells", "ngsid_reset_launch_count", "ngsid_sync", "ngsid_phase_ms", "ngsid_set_option", "ngsid_upload_reads",
           "ngsid_minimizers", "ngsid_minimizers_timed", "ngsid_get_minimizers",
           "ngsid_quality_stats", "ngsid_get_quality_stats", "ngsid_sort_scores", "ngsid_cluster", "ngsid_sg_block_align", "ngsid_sg_align_paths", "ngsid_poa_consensus", "ngsid_poa_consensus_sub",
           "ngsid_fastq_parse", "ngsid_nccl_unique_id", "ngsid_nccl_init", "ngsid_nccl_finalize", "ngsid_nccl_share", "ngsid_allgather_bytes",
           "ngsid_allreduce", "ngsid_gather_representatives", "ngsid_exchange_reads", "ngsid_append_revcomp",
           "ngsid_download_reads", "ngsid_kmer_string", "ngsid_hit_counts", "ngsid_pinned_alloc", "ngsid_pinned_free"]


class ClusterParams(ctypes.Structure):
    _fields_ = [("k", ctypes.c_int32), ("w", ctypes.c_int32), ("min_shared", ctypes.c_int32),
                ("symmetric", ctypes.c_int32), ("min_fraction", ctypes.c_double),
                ("mapped_threshold", ctypes.c_double), ("aligned_threshold", ctypes.c_double),
                ("max_gap", ctypes.c_int32 * 225), ("tile_reads", ctypes.c_int32),
                ("reserved", ctypes.c_int32 * 7)]


class PoaParams(ctypes.Structure):
    _fields_ = [("mode", ctypes.c_int32), ("match", ctypes.c_int32), ("mismatch", ctypes.c_int32),
                ("gap", ctypes.c_int32), ("trim", ctypes.c_int32), ("max_nodes", ctypes.c_int32),
                ("reserved", ctypes.c_int32 * 2)]


class ClusterStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in
                ("n_processed", "n_new_reps", "n_mapped", "n_aln_called", "n_aln_passed",
                 "n_alignments", "n_tiles", "n_chain_steps", "n_surprises", "n_map_launch_reads",
                 "align_cells")] + [("reserved", ctypes.c_int64 * 5)]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_ if n != "reserved"}


_lib = None


def load():
    """Loads libngsid.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libngsid.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback for the hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    P = ctypes.POINTER
    lib.ngsid_version.restype = i32
    lib.ngsid_ctx_create.argtypes = [i32, P(vp)]
    lib.ngsid_ctx_destroy.argtypes = [vp]
    lib.ngsid_ctx_destroy.restype = None
    lib.ngsid_last_error.argtypes = [vp]
    lib.ngsid_last_error.restype = ctypes.c_char_p
    lib.ngsid_launch_count.argtypes = [vp]
    lib.ngsid_launch_count.restype = i64
    lib.ngsid_poa_cells.argtypes = [vp]
    lib.ngsid_poa_cells.restype = i64
    lib.ngsid_reset_launch_count.argtypes = [vp]
    lib.ngsid_reset_launch_count.restype = None
    lib.ngsid_sync.argtypes = [vp]
    lib.ngsid_phase_ms.argtypes = [vp, i32]
    lib.ngsid_phase_ms.restype = ctypes.c_float
    lib.ngsid_set_option.argtypes = [vp, i32, i32]
    lib.ngsid_upload_reads.argtypes = [vp, vp, vp, vp, i64]
    lib.ngsid_minimizers.argtypes = [vp, i32, i32]
    lib.ngsid_minimizers_timed.argtypes = [vp, i32, i32, i32, P(ctypes.c_float)]
    lib.ngsid_get_minimizers.argtypes = [vp, i64, i64, vp, vp, vp, vp, i64, P(i64)]
    lib.ngsid_quality_stats.argtypes = [vp, vp, vp]
    lib.ngsid_get_quality_stats.argtypes = [vp, i64, i64, vp, vp, vp]
    lib.ngsid_sort_scores.argtypes = [vp, i32, vp, vp, vp, vp]
    lib.ngsid_cluster.argtypes = [vp, P(ClusterParams), vp, i64, vp, i64, vp, vp, vp, P(ClusterStats)]
    lib.ngsid_sg_block_align.argtypes = [vp, vp, vp, vp, vp, i64, i32, vp, vp]
    lib.ngsid_sg_align_paths.argtypes = [vp, vp, vp, vp, i64, vp, vp, i64, i32, vp, vp, vp, vp]
    lib.ngsid_poa_consensus.argtypes = [vp, P(PoaParams), i64, vp, vp, vp, vp, vp, vp, i64, vp, i64, vp, vp]
    lib.ngsid_poa_consensus_sub.argtypes = [vp, P(PoaParams), i64, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp, i64, vp, vp]
    lib.ngsid_fastq_parse.argtypes = [vp, i64, i64, vp, vp, vp, vp, vp, vp, vp, P(i64)]
    lib.ngsid_nccl_unique_id.argtypes = [vp, i64]
    lib.ngsid_nccl_init.argtypes = [vp, vp, i32, i32]
    lib.ngsid_nccl_finalize.argtypes = [vp]
    lib.ngsid_nccl_share.argtypes = [vp, vp]
    lib.ngsid_allgather_bytes.argtypes = [vp, vp, i64, vp, i64, vp]
    lib.ngsid_allreduce.argtypes = [vp, vp, i64, i32, i32]
    lib.ngsid_gather_representatives.argtypes = [vp, vp, i64, vp, vp]
    lib.ngsid_exchange_reads.argtypes = [vp, vp, vp, vp, i64, vp, vp, i64, vp]
    lib.ngsid_append_revcomp.argtypes = [vp]
    lib.ngsid_download_reads.argtypes = [vp, i64, i64, vp, vp, vp]
    lib.ngsid_kmer_string.argtypes = [vp, ctypes.c_uint32, ctypes.c_char_p, i32]
    lib.ngsid_pinned_alloc.argtypes = [P(vp), i64]
    lib.ngsid_pinned_free.argtypes = [vp]
    lib.ngsid_hit_counts.argtypes = [vp, vp, i64, vp, i64, vp, vp]
    for name in EXPORTS:
        getattr(lib, name)
    _lib = lib
    return lib


def ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class NgsidError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "libngsid error %d: %s" % (code, msg))
        self.code = code


def as_array(x, dtype):
    return np.ascontiguousarray(x, dtype=dtype)
