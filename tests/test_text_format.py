"""The native text of the loop table (cl2_loops_text / cl2_format_floats, host code): floats laid out exactly as
Python's repr() does, rows byte-identical to loops2txt on Loop objects (cLoops2/io.py:517-561 prints through str())."""
import ctypes

import numpy as np

from cloops2_b200 import _lib, hostops
from cloops2_b200.callCisLoops import cols_to_loops
from cloops2_b200.io import loops2txt, cols2txt


def test_float_text_equals_python_repr():
    _lib.build()
    L = _lib.load()
    rng = np.random.default_rng(20261020)
    v = np.concatenate([
        rng.random(100000) * 10.0 ** rng.integers(-320, 300, 100000),
        rng.integers(-10**17, 10**17, 20000).astype(np.float64),
        rng.integers(0, 10**6, 20000) / 2.0,
        np.ldexp(rng.random(50000), rng.integers(-1074, 1023, 50000)),
        np.array([0.0, -0.0, 1e16, 1e15, 9999999999999998.0, 1e-4, 9.999e-5, 1e-5, 123456789012345680.0, 5e-324, 1e-300,
                  2.2250738585072014e-308, 1.7976931348623157e308, 0.1, 0.5, 71.0, 1e22, 1e23, np.inf, -np.inf])])
    buf = ctypes.create_string_buffer(len(v) * 40)
    m = L.cl2_format_floats(v.ctypes.data_as(_lib._f64p), len(v), buf, len(buf))
    assert m > 0
    assert buf.raw[:m].decode().split("\n")[:-1] == [repr(x) for x in v.tolist()]


def _fake(rng, n):
    c = np.sort(rng.integers(0, 2 * 10**8, (n, 4)), axis=1).astype(np.int64)
    d = hostops._empty_cols()
    d.update(coords=c, ra=rng.integers(50, 3000, n), rb=rng.integers(50, 3000, n), rab=rng.integers(5, 80, n),
             anc=rng.integers(10, 500, (n, 4)).astype(np.float64))
    for k in ("P2LL", "FDR", "ES", "density", "esx", "esy", "mrabs", "mbps"):
        d[k] = rng.random(n) * 10.0 ** rng.integers(-3, 4, n)
    for k in ("hyp", "pop", "nbp", "px", "py"):
        d[k] = np.maximum(1e-300, 10.0 ** (-rng.random(n) * 320))
    if n:
        d["FDR"][:5] = 0.0
        d["ES"][:3] = 71.0
    return d


def test_table_text_equals_loops2txt(tmp_path):
    rng = np.random.default_rng(5)
    for cis, keys in ((True, [("chr1", "chr1"), ("chr2", "chr2"), ("chrX", "chrX")]), (False, [("chr1", "chr2"), ("chr3", "chrX")])):
        items = [(k, _fake(rng, n)) for k, n in zip(keys, (3000, 0, 1234))]
        loops = []
        for k, d in items:
            loops += cols_to_loops(k, d, cis=cis)
        a, b = str(tmp_path / "a.txt"), str(tmp_path / "b.txt")
        loops2txt(loops, a)
        cols2txt(items, b, cis=cis)
        assert open(a, "rb").read() == open(b, "rb").read()
