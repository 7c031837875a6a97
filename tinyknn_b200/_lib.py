"""ctypes binding of libtinyknn_b200.so (the C ABI in include/tinyknn_b200.h).

There is no CPU fallback: if the shared library is missing the import of any kernel-facing module
fails loudly, and if no CUDA device is present every compute entry point raises RuntimeError.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtinyknn_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE = 0, 1, 2, 3
ORDER_SSE, ORDER_AVX = 0, 1
DTYPE_F32, DTYPE_F64 = 0, 1
PROBE_SKIP = -(2 ** 31)

_c = ctypes
_vp, _i, _i64, _dbl = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_double

# name -> argtypes ; every function returns int status unless noted
SIGNATURES = {
    "tkb_version": [],
    "tkb_last_error": [],
    "tkb_device_count": [_c.POINTER(_c.c_int)],
    "tkb_estimate_pq_host": [_vp, _i64, _i, _vp, _vp, _i, _i],
    "tkb_query_pq_host": [_vp, _i64, _i, _i, _vp, _vp, _vp, _i, _i, _i, _vp],
    "tkb_init_heap": [_vp, _vp, _i, _i],
    "tkb_insert": [_vp, _vp, _i, _i64, _i],
    "tkb_insert_is": [_vp, _vp, _i, _i64, _i],
    "tkb_lut_build_dev": [_vp, _i, _i, _i, _vp, _vp, _i, _i, _vp, _i, _dbl, _dbl, _i, _vp, _vp, _vp, _vp, _vp],
    "tkb_estimate_dev": [_vp, _i64, _i, _vp, _i, _vp, _i64, _i, _i, _vp],
    "tkb_launch_count": [],
    "tkb_ivf_scan_dev": [_vp, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _i64, _vp, _i64, _i, _i, _vp],
    "tkb_ivf_plan_dev": [_vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i64, _vp],
    "tkb_ivf_plan_push_dev": [_vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i64, _vp],
    "tkb_ivf_plan_pull_owner_dev": [_vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _i64, _vp, _vp, _vp, _i64, _vp],
    "tkb_ivf_plan_pull_home_dev": [_vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i64, _vp],
    "tkb_ivf_pull_minima_dev": [_vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp],
    "tkb_ivf_replay_fresh_pull_dev": [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp],
    "tkb_peer_alloc": [_i64, _c.POINTER(_vp), _vp],
    "tkb_peer_open": [_vp, _c.POINTER(_vp)],
    "tkb_peer_close": [_vp],
    "tkb_peer_free": [_vp],
    "tkb_encode_dev": [_vp, _i, _i64, _i, _vp, _i64, _vp, _vp, _i, _i, _vp, _i, _vp, _vp],
    "tkb_assign_dev": [_vp, _i, _i64, _i, _vp, _i, _vp, _vp, _i, _vp, _vp, _i64, _vp],
    "tkb_kmeans_workspace": [_i64, _i, _i, _c.POINTER(_c.c_int64)],
    "tkb_kmeans_dev": [_vp, _i64, _i, _i, _vp, _i, _dbl, _vp, _c.POINTER(_c.c_int), _vp, _i64, _vp],
    "tkb_kmeans_pq_dev": [_vp, _i64, _i, _i, _vp, _i, _dbl, _vp, _i64, _vp],
    "tkb_codes_to_native_dev": [_vp, _i64, _i, _vp, _vp],
    "tkb_codes_from_native_dev": [_vp, _i64, _i, _vp, _vp],
    "tkb_estimate_native_dev": [_vp, _i64, _i, _vp, _i, _vp, _i64, _i, _i, _vp, _i64, _vp],
    "tkb_ivf_scan_native_dev": [_vp, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _i64, _vp, _i64, _i, _i, _vp, _i64, _vp],
    "tkb_heap_fill_dev": [_vp, _vp, _i64, _i, _vp],
    "tkb_replay_dev": [_vp, _i64, _i64, _i, _vp, _vp, _i, _i, _i, _vp, _vp],
    "tkb_ivf_replay_dev": [_vp, _i64, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp],
    "tkb_replay_fresh_dev": [_vp, _i64, _i64, _i, _vp, _vp, _i, _i, _i, _vp],
    "tkb_ivf_replay_fresh_dev": [_vp, _i64, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp],
    "tkb_ivf_scan_tc_supported": [],
    "tkb_ivf_scan_tc_workspace": [_i, _i, _i, _c.POINTER(_c.c_int64)],
    "tkb_ivf_scan_tc_dev": [_vp, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _i64, _vp, _i64, _vp],
    "tkb_ivf_scan_native_cm_dev": [_vp, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _vp, _i64, _i, _i, _vp, _i64, _vp],
    "tkb_ivf_scan_native_push_cm_dev": [_vp, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _i64, _i, _i, _vp, _i64, _vp],
    "tkb_ivf_replay_fresh_cm_dev": [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp],
    "tkb_gather_dists_dev": [_vp, _i, _i64, _i, _vp, _vp, _i, _i, _vp, _vp],
    "tkb_select_probes_dev": [_vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "tkb_coarse_probes_dev": [_vp, _i64, _i, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "tkb_select_topk_dev": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "tkb_ivf_query_fused_workspace": [_i, _i, _i, _i, _i, _i, _i64, _c.POINTER(_c.c_int64)],
    "tkb_ivf_query_fused_dev": [_vp, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _i64, _i, _vp, _i, _i, _i, _i64,
                                _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp],
}


class TinyKnnError(RuntimeError):
    pass


def _load():
    from . import build as _build
    if _build.needs_build():
        # sources changed (or first use): rebuild in-tree when a compiler is around. One process at a time (torchrun starts
        # every rank at once: they would all write the same object files), the others wait for the lock and find it built.
        import fcntl
        lock = open(LIB_PATH + ".lock", "w")
        try:
            fcntl.flock(lock, fcntl.LOCK_EX)
            if _build.needs_build():
                try:
                    _build.build(force=True)
                except Exception as e:                               # noqa: BLE001
                    # never load a library built from OTHER sources: a changed C signature under an unchanged symbol name would
                    # get mis-typed ctypes arguments. TKB_ALLOW_STALE_LIB=1 overrides (no compiler on this machine, sources touched).
                    if not os.path.exists(LIB_PATH) or os.environ.get("TKB_ALLOW_STALE_LIB", "0") == "0":
                        raise ImportError(
                            "tinyknn_b200: %s is missing or older than its sources and could not be built (%s). Build it with "
                            "`python -m tinyknn_b200.build` (nvcc, sm_100a). There is no CPU fallback." % (LIB_PATH, e))
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
            lock.close()
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export the symbol
        fn.argtypes = argtypes
        fn.restype = {"tkb_last_error": _c.c_char_p, "tkb_launch_count": _c.c_longlong}.get(name, _c.c_int)
    return lib


lib = _load()


def last_error():
    msg = lib.tkb_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


n_calls = 0          # C-ABI calls that returned OK
PLAN_SEND, PLAN_RECV, PLAN_PUSH = 0, 1, 2


def launch_count():
    """CUDA kernels launched by the library so far (counted where they are launched, in C)."""
    return int(lib.tkb_launch_count())


def check(rc):
    global n_calls
    if rc == OK:
        n_calls += 1
        return
    msg = last_error()
    if rc == ERR_INVALID:
        raise ValueError("tinyknn_b200: " + msg)
    raise TinyKnnError("tinyknn_b200: " + msg)


def device_count():
    n = _c.c_int(0)
    check(lib.tkb_device_count(_c.byref(n)))
    return n.value
