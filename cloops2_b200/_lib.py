"""
ctypes binding of libcl2.so (include/cl2.h) -- the only compute path of this package.

There is deliberately NO CPU fallback: if the CUDA library is missing or no GPU is visible, every entry point
raises (Cl2Error / OSError).  The library is built in-tree by `build()` (nvcc, sm_100a) so it travels with the
repository snapshot to the GPU box.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.path.join(_HERE, "libcl2.so")
SOURCES = ["cl2_api.cu", "cl2_sort.cu", "cl2_block.cu", "cl2_cdbscan.cu", "cl2_count.cu", "cl2_driver.cu", "cl2_comm.cu", "cl2_host.cu",
           "cl2_pre.cu"]
HEADERS = ["cl2_common.cuh", "cl2_points.cuh", "cl2_stats.cuh", os.path.join("..", "..", "include", "cl2.h")]
# -fmad=false: no FMA contraction, so the fp64 statistics give the same bits in every kernel they are inlined into
# (and the same bits as the host build of cl2_stats.cuh); the path is integer / bandwidth work, nothing is lost.
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC"]

MODE_CIS, MODE_TRANS = 0, 1


class Cl2Error(RuntimeError):
    pass


def source_hash():
    """sha256 over the CUDA sources, headers and compiler flags of libcl2.so (first 16 hex digits)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in SOURCES + HEADERS:
        with open(os.path.join(_CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]


HASH_PATH = SO_PATH + ".hash"


def _stale():
    """Does the shipped libcl2.so come from these sources?  Decided by content (the hash written beside the library
    when it was built, also compiled into cl2_version()), not by file times: a snapshot copy does not keep them."""
    if not os.path.exists(SO_PATH) or not os.path.exists(HASH_PATH):
        return True
    try:
        return open(HASH_PATH).read().strip() != source_hash()
    except OSError:
        return True


def build(force=False, verbose=False):
    """Compile libcl2.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    if not force and not _stale():
        return SO_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    digest = source_hash()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(_CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ['-DCL2_SRC_HASH="%s"' % digest, "-c", "-o", obj, os.path.join(_CSRC, src)]
        if verbose:
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise Cl2Error("nvcc failed: %s\n%s" % (" ".join(cmd), out.decode(errors="replace")))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", SO_PATH] + objs + ["-lcudart", "-ldl"]
    subprocess.check_call(cmd)
    with open(HASH_PATH, "w") as fh:
        fh.write(digest + "\n")
    return SO_PATH


_i64p = ctypes.POINTER(ctypes.c_int64)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_f64p = ctypes.POINTER(ctypes.c_double)
_vp = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/cl2.h one to one (tests/test_abi.py checks the two agree)
SIGNATURES = {
    "cl2_version": (ctypes.c_char_p, []),
    "cl2_ctx_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_vp)]),
    "cl2_ctx_destroy": (None, [_vp]),
    "cl2_last_error": (ctypes.c_char_p, [_vp]),
    "cl2_ctx_sync": (ctypes.c_int, [_vp]),
    "cl2_host_alloc": (ctypes.c_int, [_vp, ctypes.c_int64, ctypes.POINTER(_vp)]),
    "cl2_host_free": (None, [_vp, _vp]),
    "cl2_points_upload": (ctypes.c_int, [_vp, _i64p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(_vp)]),
    "cl2_points_count": (ctypes.c_int64, [_vp]),
    "cl2_points_extent": (ctypes.c_int, [_vp, _vp, _i64p]),
    "cl2_points_download": (ctypes.c_int, [_vp, _vp, _i64p]),
    "cl2_points_free": (None, [_vp, _vp]),
    "cl2_points_drop_cache": (None, [_vp, _vp]),
    "cl2_points_sweep_hint": (ctypes.c_int, [_vp, _vp, ctypes.c_int32]),
    "cl2_points_index_hint": (ctypes.c_int, [_vp, _vp, ctypes.c_int32]),
    "cl2_points_unique": (ctypes.c_int, [_vp, _vp, _i64p, ctypes.POINTER(ctypes.c_int64)]),
    "cl2_set_sync_mode": (None, [ctypes.c_int32]),
    "cl2_bedpe_create": (ctypes.c_int, [ctypes.POINTER(_vp)]),
    "cl2_bedpe_destroy": (None, [_vp]),
    "cl2_bedpe_parse": (ctypes.c_int64, [_vp, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int64, _i32p, _i64p, _i32p,
                                         ctypes.POINTER(ctypes.c_int64)]),
    "cl2_pairs_parse": (ctypes.c_int64, [_vp, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int64, _i32p, _i64p,
                                         ctypes.POINTER(ctypes.c_int64)]),
    "cl2_bedpe_names": (ctypes.c_char_p, [_vp, ctypes.POINTER(ctypes.c_int32)]),
    "cl2_block_dbscan": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int32, _i32p, _i32p]),
    "cl2_cdbscan2": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int32, _i32p, _i32p]),
    "cl2_cluster_bbox": (ctypes.c_int, [_vp, _vp, _i64p, _i64p]),
    "cl2_block_noise": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int32, _u8p, _i32p]),
    "cl2_index_build": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(_vp)]),
    "cl2_index_free": (None, [_vp, _vp]),
    "cl2_loop_counts": (ctypes.c_int, [_vp, _vp, _i64p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                       _i64p, _i64p, _i64p, _i64p, _i64p]),
    "cl2_loop_core": (ctypes.c_int, [_vp, _vp, _i64p, ctypes.c_int64, ctypes.c_int32, _i64p, _f64p]),
    "cl2_loop_tests": (ctypes.c_int, [_vp, _vp, _i64p, _i64p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                      ctypes.c_double, _f64p]),
    "cl2_est_loops": (ctypes.c_int, [_vp, _vp, _i64p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                     ctypes.c_int32, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                     _i64p, _i64p, _i64p, _f64p, _f64p]),
    "cl2_set_pvalue_margin": (ctypes.c_int, [_vp, ctypes.c_double]),
    "cl2_call_loops": (ctypes.c_int, [_vp, _vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _i64p, _i32p, ctypes.c_int64,
                                      ctypes.c_int32, ctypes.c_int32, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                      ctypes.c_int32, ctypes.POINTER(_vp)]),
    "cl2_loops_count": (ctypes.c_int64, [_vp]),
    "cl2_loops_info": (ctypes.c_int, [_vp, _i64p, _i64p, _i64p, _f64p]),
    "cl2_loops_read": (ctypes.c_int, [_vp, _i64p, _i32p, _i64p, _f64p, _f64p]),
    "cl2_loops_free": (None, [_vp]),
    "cl2_comm_unique_id": (ctypes.c_int, [_u8p]),
    "cl2_comm_init": (ctypes.c_int, [_vp, _u8p, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(_vp)]),
    "cl2_stats_allreduce": (ctypes.c_int, [_vp, _vp, _i64p, ctypes.c_int32]),
    "cl2_comm_destroy": (None, [_vp]),
    "cl2_sel_sig_loops": (ctypes.c_int, [_i64p, _f64p, ctypes.c_int64, _i64p, _i64p]),
    "cl2_loops_text": (ctypes.c_int64, [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, _i64p,
                                        _i64p, _i64p, _i64p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p,
                                        ctypes.c_char_p, ctypes.c_int64]),
    "cl2_format_floats": (ctypes.c_int64, [_f64p, ctypes.c_int64, ctypes.c_char_p, ctypes.c_int64]),
    "cl2_stat_poisson_sf": (ctypes.c_double, [ctypes.c_double, ctypes.c_double]),
    "cl2_stat_binom_sf": (ctypes.c_double, [ctypes.c_double, ctypes.c_double, ctypes.c_double]),
    "cl2_stat_hypergeom_sf": (ctypes.c_double, [ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double]),
    "cl2_launch_count": (ctypes.c_int64, [_vp]),
    "cl2_launch_count_reset": (None, [_vp]),
    "cl2_timing_enable": (ctypes.c_int, [_vp, ctypes.c_int]),
    "cl2_transfer_bytes": (None, [_vp, _i64p, _i64p]),
    "cl2_timing_read": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_char_p, _f64p, _i64p, _f64p]),
    "cl2_timing_reset": (None, [_vp]),
}

_lib = None


def load():
    """dlopen libcl2.so and set the prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise Cl2Error("libcl2.so not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'`" % SO_PATH)
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        if os.environ.get("CL2_BLOCKING_SYNC", "0") not in ("0", ""):
            L.cl2_set_sync_mode(1)
        _lib = L
    return _lib


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(t)


class Context:
    """One GPU (cl2_ctx): stream, scratch pool, kernel timers."""

    def __init__(self, device=0):
        self.lib = load()
        h = _vp()
        rc = self.lib.cl2_ctx_create(int(device), ctypes.byref(h))
        if rc != 0 or not h:
            raise Cl2Error("cl2_ctx_create(device=%d) failed rc=%d (no CUDA device visible?)" % (device, rc))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.cl2_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc, what):
        if rc != 0:
            msg = self.lib.cl2_last_error(self.h)
            raise Cl2Error("%s failed rc=%d: %s" % (what, rc, msg.decode(errors="replace") if msg else ""))

    def sync(self):
        self.check(self.lib.cl2_ctx_sync(self.h), "cl2_ctx_sync")

    def set_pvalue_margin(self, margin):
        """Widen the device-side p-value cuts by 1 + margin (cl2_set_pvalue_margin)."""
        self.check(self.lib.cl2_set_pvalue_margin(self.h, float(margin)), "cl2_set_pvalue_margin")

    # ---- pinned host staging
    def pinned_empty(self, shape, dtype=np.int64):
        """numpy array backed by page-locked memory (freed with the returned array's owner)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) if len(np.shape(shape)) else int(shape)
        nbytes = max(1, n * dtype.itemsize)
        p = _vp()
        self.check(self.lib.cl2_host_alloc(self.h, nbytes, ctypes.byref(p)), "cl2_host_alloc")
        buf = (ctypes.c_char * nbytes).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
        owner = _PinnedOwner(self, p)
        arr = arr.view(_PinnedArray)
        arr._owner = owner
        return arr

    def staging(self, shape, dtype=np.int64):
        """A view of this context's grow-only page-locked staging buffer (one per worker; reused between files)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape))
        nbytes = max(1, n * dtype.itemsize)
        buf = getattr(self, "_staging", None)
        if buf is None or buf.nbytes < nbytes:
            self._staging = None
            buf = self._staging = self.pinned_empty(nbytes + nbytes // 8, np.uint8)
        return buf[:n * dtype.itemsize].view(dtype).reshape(shape)

    # ---- measurement
    def launch_count(self):
        return int(self.lib.cl2_launch_count(self.h))

    def launch_count_reset(self):
        self.lib.cl2_launch_count_reset(self.h)

    def transfer_bytes(self):
        a, b = ctypes.c_int64(0), ctypes.c_int64(0)
        self.lib.cl2_transfer_bytes(self.h, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    def timing_enable(self, on=True):
        self.check(self.lib.cl2_timing_enable(self.h, 1 if on else 0), "cl2_timing_enable")

    def timing_reset(self):
        self.lib.cl2_timing_reset(self.h)

    def timing_read(self):
        k = 32
        names = ctypes.create_string_buffer(32 * k)
        ms = (ctypes.c_double * k)()
        ln = (ctypes.c_int64 * k)()
        by = (ctypes.c_double * k)()
        m = self.lib.cl2_timing_read(self.h, k, names, ms, ln, by)
        out = {}
        for i in range(m):
            nm = names.raw[32 * i:32 * i + 32].split(b"\0")[0].decode()
            out[nm] = {"ms": ms[i], "launches": int(ln[i]), "bytes": by[i]}
        return out


class _PinnedOwner:
    def __init__(self, ctx, p):
        self.ctx, self.p = ctx, p

    def __del__(self):
        try:
            if self.p and self.ctx.h:
                self.ctx.lib.cl2_host_free(self.ctx.h, self.p)
        except Exception:
            pass
        self.p = None


class _PinnedArray(np.ndarray):
    _owner = None

    def __array_finalize__(self, obj):
        if obj is not None and getattr(obj, "_owner", None) is not None:
            self._owner = obj._owner


class Points:
    """One .ixy matrix resident in HBM (cl2_points); the reference's `mat` after parseIxy (io.py:274-291)."""

    def __init__(self, ctx, xy, cut=0, mcut=-1):
        self.ctx = ctx
        xy = np.ascontiguousarray(xy, dtype=np.int64)
        if xy.ndim != 2 or xy.shape[1] != 2:
            raise ValueError("xy must be int64[N,2]")
        h = _vp()
        ctx.check(ctx.lib.cl2_points_upload(ctx.h, _ptr(xy, _i64p), xy.shape[0], int(cut), int(mcut), ctypes.byref(h)),
                  "cl2_points_upload")
        self.h = h
        self.n = int(ctx.lib.cl2_points_count(h))
        self.ncl = 0

    def close(self):
        if getattr(self, "h", None) and self.ctx.h and not getattr(self, "_view", False):
            self.ctx.lib.cl2_points_free(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def on(self, ctx):
        """The same resident points driven through another context (stream) of the same GPU; the view does not own
        them.  The C ABI takes the context with every call, so a file uploaded once can be processed by whichever
        worker is free."""
        v = object.__new__(Points)
        v.ctx, v.h, v.n, v.ncl, v._view = ctx, self.h, self.n, 0, True
        return v

    def drop_cache(self):
        """Free the cached eps grid (the points stay resident)."""
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.cl2_points_drop_cache(self.ctx.h, self.h)

    def unique(self):
        """Rows without exact duplicates, ascending by (X, Y): int64[m,2] (cl2_points_unique, the dedup of `pre`)."""
        out = np.empty((self.n, 2), dtype=np.int64)
        m = ctypes.c_int64(0)
        self.ctx.check(self.ctx.lib.cl2_points_unique(self.ctx.h, self.h, _ptr(out, _i64p), ctypes.byref(m)),
                       "cl2_points_unique")
        return out[:m.value]

    def sweep_hint(self, enable=True):
        """Announce blockDBSCAN runs with decreasing eps: later runs then skip the rows an earlier run proved to be
        noise for them (cl2_points_sweep_hint); results do not change.  enable: False/0 off, True/1 remember and
        use, 2 use only."""
        self.ctx.check(self.ctx.lib.cl2_points_sweep_hint(self.ctx.h, self.h, int(enable)), "cl2_points_sweep_hint")

    def index_hint(self, enable=True):
        """Announce that an index build follows the next blockDBSCAN runs: the first run over every row then leaves the
        x-order of the points behind for it (cl2_points_index_hint); the index is the same either way."""
        self.ctx.check(self.ctx.lib.cl2_points_index_hint(self.ctx.h, self.h, int(bool(enable))), "cl2_points_index_hint")

    def extent(self):
        ext = np.zeros(4, dtype=np.int64)
        self.ctx.check(self.ctx.lib.cl2_points_extent(self.ctx.h, self.h, _ptr(ext, _i64p)), "cl2_points_extent")
        return ext

    def download(self):
        out = np.empty((self.n, 2), dtype=np.int64)
        self.ctx.check(self.ctx.lib.cl2_points_download(self.ctx.h, self.h, _ptr(out, _i64p)), "cl2_points_download")
        return out

    def _dbscan(self, fn, name, eps, minPts, want_labels):
        labels = np.empty(self.n, dtype=np.int32) if want_labels else None
        ncl = ctypes.c_int32(0)
        self.ctx.check(fn(self.ctx.h, self.h, int(eps), int(minPts), _ptr(labels, _i32p), ctypes.byref(ncl)), name)
        self.ncl = ncl.value
        return labels, ncl.value

    def block_dbscan(self, eps, minPts, want_labels=True):
        return self._dbscan(self.ctx.lib.cl2_block_dbscan, "cl2_block_dbscan", eps, minPts, want_labels)

    def cdbscan2(self, eps, minPts, want_labels=True):
        return self._dbscan(self.ctx.lib.cl2_cdbscan2, "cl2_cdbscan2", eps, minPts, want_labels)

    def cluster_bbox(self):
        bbox = np.empty((self.ncl, 4), dtype=np.int64)
        size = np.empty(self.ncl, dtype=np.int64)
        if self.ncl:
            self.ctx.check(self.ctx.lib.cl2_cluster_bbox(self.ctx.h, self.h, _ptr(bbox, _i64p), _ptr(size, _i64p)),
                           "cl2_cluster_bbox")
        return bbox, size

    def call_loops(self, passes, mode=MODE_CIS, hic=False, cut=0, min_pts_test=5, win=5, countDiffCut=20, pseudo=1,
                   peakPcut=1e-5, want_unique=False):
        """The whole per-file work unit in one native call (cl2_call_loops): every clustering run of `passes`
        [(eps, minPts), ...], candidate filters, estLoopSig, de-duplication.  Returns a dict: coords int64[K,4],
        pass_id int32[K], core int64[K,4], hyp float64[K], stats float64[K,13] of the surviving loops in candidate
        order, plus pass_cand / pass_pets per run, totals and the host seconds per stage."""
        npass = len(passes)
        eps = np.array([int(e) for e, _ in passes], dtype=np.int64)
        mps = np.array([int(m) for _, m in passes], dtype=np.int32)
        h = _vp()
        lib, ctx = self.ctx.lib, self.ctx
        ctx.check(lib.cl2_call_loops(ctx.h, self.h, int(mode), 1 if hic else 0, npass, _ptr(eps, _i64p), _ptr(mps, _i32p),
                                     int(cut), int(min_pts_test), int(win), float(countDiffCut), float(pseudo),
                                     float(peakPcut), 1 if want_unique else 0, ctypes.byref(h)), "cl2_call_loops")
        try:
            k = int(lib.cl2_loops_count(h))
            out = {"coords": np.empty((k, 4), dtype=np.int64), "pass_id": np.empty(k, dtype=np.int32),
                   "core": np.empty((k, 4), dtype=np.int64), "hyp": np.empty(k, dtype=np.float64),
                   "stats": np.empty((k, 13), dtype=np.float64), "pass_cand": np.zeros(npass, dtype=np.int64),
                   "pass_pets": np.zeros(npass, dtype=np.int64), "totals": np.zeros(8, dtype=np.int64),
                   "seconds": np.zeros(6, dtype=np.float64)}
            ctx.check(lib.cl2_loops_info(h, _ptr(out["pass_cand"], _i64p), _ptr(out["pass_pets"], _i64p),
                                         _ptr(out["totals"], _i64p), _ptr(out["seconds"], _f64p)), "cl2_loops_info")
            ctx.check(lib.cl2_loops_read(h, _ptr(out["coords"], _i64p), _ptr(out["pass_id"], _i32p), _ptr(out["core"], _i64p),
                                         _ptr(out["hyp"], _f64p), _ptr(out["stats"], _f64p)), "cl2_loops_read")
        finally:
            lib.cl2_loops_free(h)
        return out

    def block_noise_keep(self, eps, minPts, with_cell_order=False):
        """keep[n] of blockDBSCAN's removeNoiseGrids; with_cell_order also returns, per row, the first row id of its
        cell (the reference's dict order of the cells)."""
        keep = np.zeros(self.n, dtype=np.uint8)
        first = np.zeros(self.n, dtype=np.int32) if with_cell_order else None
        self.ctx.check(self.ctx.lib.cl2_block_noise(self.ctx.h, self.h, int(eps), int(minPts), _ptr(keep, _u8p),
                                                    _ptr(first, _i32p)), "cl2_block_noise")
        return (keep.astype(bool), first) if with_cell_order else keep.astype(bool)


class Index:
    """Device XY index (cl2_index) = cLoops2/ds.py:141-198."""

    def __init__(self, pts):
        self.ctx = pts.ctx
        self.n = pts.n
        h = _vp()
        self.ctx.check(self.ctx.lib.cl2_index_build(self.ctx.h, pts.h, ctypes.byref(h)), "cl2_index_build")
        self.h = h

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.cl2_index_free(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def loop_counts(self, loops, win=5, mode=MODE_CIS, core=True, windows=True, anchors=True):
        loops = np.ascontiguousarray(loops, dtype=np.int64).reshape(-1, 4)
        L = loops.shape[0]
        W = 2 * win
        o_core = np.zeros((L, 4), dtype=np.int64) if core else None
        o_na = np.zeros((L, W), dtype=np.int64) if windows else None
        o_nb = np.zeros((L, W), dtype=np.int64) if windows else None
        o_nab = np.zeros((L, W * W), dtype=np.int64) if windows else None
        o_anc = np.zeros((L, 4), dtype=np.int64) if anchors else None
        self.ctx.check(self.ctx.lib.cl2_loop_counts(self.ctx.h, self.h, _ptr(loops, _i64p), L, win, mode,
                                                    _ptr(o_core, _i64p), _ptr(o_na, _i64p), _ptr(o_nb, _i64p),
                                                    _ptr(o_nab, _i64p), _ptr(o_anc, _i64p)), "cl2_loop_counts")
        return o_core, o_na, o_nb, o_nab, o_anc


    def loop_core(self, loops, mode=MODE_CIS):
        """(core int64[L,4], hyp float64[L]) -- phase 1 of the fused counting + statistics."""
        loops = np.ascontiguousarray(loops, dtype=np.int64).reshape(-1, 4)
        L = loops.shape[0]
        core = np.zeros((L, 4), dtype=np.int64)
        hyp = np.zeros(L, dtype=np.float64)
        self.ctx.check(self.ctx.lib.cl2_loop_core(self.ctx.h, self.h, _ptr(loops, _i64p), L, mode, _ptr(core, _i64p),
                                                  _ptr(hyp, _f64p)), "cl2_loop_core")
        return core, hyp

    def est_loops(self, loops, rpb, minPts, mode=MODE_CIS, win=5, hic=False, countDiffCut=20, pseudo=1, peakPcut=1e-5):
        """Counting + statistics + the reference's filter cascade on the device; returns only the survivors:
        (sel int64[K] indices into loops, core int64[K,4], hyp float64[K], stats float64[K,13])."""
        loops = np.ascontiguousarray(loops, dtype=np.int64).reshape(-1, 4)
        L = loops.shape[0]
        sel = np.empty(L, dtype=np.int64)
        core = np.empty((L, 4), dtype=np.int64)
        hyp = np.empty(L, dtype=np.float64)
        stats = np.empty((L, 13), dtype=np.float64)
        n = ctypes.c_int64(0)
        self.ctx.check(self.ctx.lib.cl2_est_loops(self.ctx.h, self.h, _ptr(loops, _i64p), L, win, mode, int(minPts),
                                                  1 if hic else 0, float(countDiffCut), float(pseudo), float(peakPcut),
                                                  float(rpb), ctypes.byref(n), _ptr(sel, _i64p), _ptr(core, _i64p),
                                                  _ptr(hyp, _f64p), _ptr(stats, _f64p)), "cl2_est_loops")
        k = n.value
        return sel[:k], core[:k], hyp[:k], stats[:k]

    def loop_tests(self, loops, core, rpb, win=5, mode=MODE_CIS):
        """stats float64[L,13] -- phase 2 (see include/cl2.h)."""
        loops = np.ascontiguousarray(loops, dtype=np.int64).reshape(-1, 4)
        core = np.ascontiguousarray(core, dtype=np.int64).reshape(-1, 4)
        L = loops.shape[0]
        stats = np.zeros((L, 13), dtype=np.float64)
        self.ctx.check(self.ctx.lib.cl2_loop_tests(self.ctx.h, self.h, _ptr(loops, _i64p), _ptr(core, _i64p), L, win, mode,
                                                   float(rpb), _ptr(stats, _f64p)), "cl2_loop_tests")
        return stats


_default_ctx = {}


def default_context(device=None):
    """Process-wide context for `device` (LOCAL_RANK when launched by torchrun, else 0)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
