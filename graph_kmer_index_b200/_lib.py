"""ctypes binding of libgki.so (include/gki.h) + the in-tree nvcc build recipe.

There is no CPU fallback: if the library cannot be loaded, or a call fails (no CUDA device, bad
argument, ...), a GkiError is raised.
"""
import ctypes
import glob
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")
LIB_PATH = os.path.join(_PKG, "libgki.so")
SOURCES = ["runtime.cu", "scan.cu", "hash.cu", "index.cu", "count.cu", "build.cu", "synth.cu", "finder.cu", "comm.cu", "ingest.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-lpthread", "-ldl"]


class GkiError(RuntimeError):
    pass


def _sources():
    return [os.path.join(_CSRC, s) for s in SOURCES if os.path.exists(os.path.join(_CSRC, s))]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = _sources() + glob.glob(os.path.join(_CSRC, "*.cuh")) + glob.glob(os.path.join(_CSRC, "*.h")) + [os.path.join(_PKG, "..", "include", "gki.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into graph_kmer_index_b200/libgki.so (nvcc cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()      # linked beside the target and renamed: a reader never sees a half-written library
    knobs = ["-DGKI_EXPERIMENT_KNOBS"] if os.environ.get("GKI_BUILD_EXPERIMENT_KNOBS") == "1" else []   # profiles/ sweeps only (csrc/common.cuh)
    cmd = [nvcc] + NVCC_FLAGS + knobs + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + _sources()
    try:
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    except FileNotFoundError as e:
        raise GkiError("nvcc not found; libgki.so must be prebuilt in-tree (%s)" % e)
    if out.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise GkiError("nvcc failed:\n" + out.stdout)
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(out.stdout)
    return LIB_PATH


c_i32, c_i64, c_u64, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_void_p

# name -> argtypes (every function returns int status except where noted)
_SIGNATURES = {
    "gki_device_count": [ctypes.POINTER(ctypes.c_int)],
    "gki_set_device": [ctypes.c_int],
    "gki_encode_bases": [c_vp, c_i64, c_vp, c_vp],
    "gki_hash_reads": [c_vp, c_i64, c_i32, c_i64, c_i32, c_vp, c_vp, c_vp],
    "gki_hash_reads_ragged": [c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp],
    "gki_revcomp_hashes": [c_vp, c_i64, c_i32, c_vp, c_vp],
    "gki_complement_hashes": [c_vp, c_i64, c_i32, c_vp, c_vp],
    "gki_hashes_to_bases": [c_vp, c_i64, c_i32, c_vp, c_vp],
    "gki_index_build": [c_vp, c_vp, c_vp, c_vp, c_i64, c_u64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "gki_partition_by_bucket_range": [c_vp, c_i64, c_u64, c_i32, c_vp, c_vp, c_vp],
    "gki_index_build_range": [c_vp, c_vp, c_vp, c_vp, c_i64, c_u64, c_u64, c_u64, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "gki_partition_pack": [c_vp, c_vp, c_vp, c_vp, c_i64, c_u64, c_i32, c_vp, c_vp, c_vp],
    "gki_index_build_records": [c_vp, c_i64, c_u64, c_u64, c_u64, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "gki_group_by_key": [c_vp, c_i32, c_i64, c_u64, c_i32, c_vp, c_vp, c_vp, c_vp],
    "gki_gather": [c_vp, c_i32, c_vp, c_i64, c_vp, c_vp],
    "gki_mark_non_first_occurrences": [c_vp, c_i64, c_vp, c_vp],
    "gki_index_create": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_u64, c_i32, ctypes.POINTER(c_vp), c_vp],
    "gki_index_destroy": [c_vp],
    "gki_index_info": [c_vp, ctypes.POINTER(c_i64), ctypes.POINTER(c_u64), ctypes.POINTER(c_i64),
                       ctypes.POINTER(c_i64), ctypes.POINTER(c_i32), ctypes.POINTER(c_i64)],
    "gki_prepare_counting": [c_vp, c_i32, c_vp],
    "gki_pack_reads": [c_vp, c_i64, c_i32, c_i64, c_vp, c_vp, c_i64, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), c_i32, c_i32],
    "gki_count_packed_reads": [c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp],
    "gki_host_read_bandwidth": [c_vp, c_i64, c_i32, ctypes.POINTER(ctypes.c_double)],
    "gki_fastx_open": [ctypes.c_char_p, ctypes.POINTER(c_vp), ctypes.POINTER(c_i64), ctypes.POINTER(c_i32), ctypes.POINTER(c_i32)],
    "gki_fastx_lines": [c_vp, c_vp, c_vp],
    "gki_count_fastx": [c_vp, c_vp, c_i32, c_i32, ctypes.POINTER(c_i64), c_vp],
    "gki_fastx_close": [c_vp],
    "gki_reset_counts": [c_vp, c_vp],
    "gki_count_kmers": [c_vp, c_vp, c_i64, c_vp],
    "gki_count_reads": [c_vp, c_vp, c_i64, c_i32, c_i64, c_i32, c_i32, c_vp],
    "gki_node_counts": [c_vp, c_vp, c_i64, c_i32, c_vp],
    "gki_entry_counts": [c_vp, c_vp, c_vp],
    "gki_map_kmers": [c_vp, c_vp, c_i64, c_vp, c_i64, c_i32, c_i32, c_vp],
    "gki_has_kmers": [c_vp, c_vp, c_i64, c_vp, c_i32, c_vp],
    "gki_lookup_hits": [c_vp, c_vp, c_i64, c_i32, c_i64, c_i32, c_vp, c_i64, ctypes.POINTER(c_i64), c_vp],
    "gki_lookup_entries": [c_vp, c_vp, c_i64, c_i32, c_i64, c_i32, c_vp, c_vp, c_i64, ctypes.POINTER(c_i64), c_vp],
    "gki_query_counts": [c_vp, c_vp, c_i64, c_vp, c_vp],
    "gki_synth_genome": [c_vp, c_i64, c_vp],
    "gki_synth_flat_kmers": [c_vp, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp],
    "gki_synth_reads": [c_vp, c_i64, c_i64, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp],
    "gki_calibrate_random_gather": [c_i64, c_i64, c_i32, ctypes.POINTER(ctypes.c_float)],
    "gki_nccl_unique_id": [c_vp],
    "gki_nccl_comm_create": [c_vp, c_i32, c_i32, ctypes.POINTER(c_vp)],
    "gki_nccl_comm_destroy": [c_vp],
    "gki_allreduce_counts": [c_vp, c_vp, c_i64, c_i32, c_vp],
    "gki_release_scratch": [],
    "gki_calibrate_scatter": [c_i64, c_i32, c_i64, ctypes.POINTER(ctypes.c_float)],
    "gki_calibrate_store_groups": [c_i64, c_i32, c_i64, c_i32, ctypes.POINTER(ctypes.c_float)],
    "gki_calibrate_copy": [c_i64, ctypes.POINTER(ctypes.c_float)],
    "gki_critical_paths": [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i32, c_i32, c_vp, c_vp, c_i64, ctypes.POINTER(c_i64), c_vp],
    "gki_finder_prepare": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_i32, c_i32, c_i32,
                           c_i32, c_i64, ctypes.POINTER(c_vp), ctypes.POINTER(c_i64), c_vp],
    "gki_finder_fill": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "gki_finder_destroy": [c_vp],
}
EXPORTED = [n for n in _SIGNATURES] + ["gki_last_error", "gki_version", "gki_launch_count"]

GKI_BUILD_SKIP_FREQUENCIES = 1
GKI_COUNTS_WRAP_UINT16 = 1
GKI_PROBE_SKIP_BUCKET0 = 1
GKI_COUNTS_FLOAT64, GKI_COUNTS_UINT64 = 0, 1

_lib = None


def load():
    """Load (building first if the in-tree .so is stale and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    if needs_build():
        build()
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:
        raise GkiError("cannot load %s: %s -- the CUDA extension is required, there is no CPU fallback" % (LIB_PATH, e))
    for name, argtypes in _SIGNATURES.items():
        if argtypes is None or not hasattr(lib, name):
            continue
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    lib.gki_last_error.restype = ctypes.c_char_p
    lib.gki_version.restype = ctypes.c_int
    lib.gki_launch_count.restype = ctypes.c_int64
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise GkiError("libgki error %d: %s" % (status, load().gki_last_error().decode(errors="replace")))


def call(name, *args):
    check(getattr(load(), name)(*args))


def launch_count():
    return int(load().gki_launch_count())


def ptr(a):
    """Pointer of a numpy array (host) or torch tensor (host or device); None -> NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags.c_contiguous, "array must be C-contiguous"
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous(), "tensor must be contiguous"
        return a.data_ptr()
    if isinstance(a, int):
        return a
    raise TypeError("cannot take a pointer of %r" % type(a))


def current_stream():
    """cudaStream_t of torch's current stream if torch is already imported and CUDA is up, else the default stream."""
    import sys
    torch = sys.modules.get("torch")
    if torch is not None and torch.cuda.is_available() and torch.cuda.is_initialized():
        return torch.cuda.current_stream().cuda_stream
    return None
