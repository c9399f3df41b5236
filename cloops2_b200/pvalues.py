"""
The scipy tail probabilities of the exact mode, argument for argument as the reference calls them
(callCisLoops.py:246, 310, 329, 332; callTransLoops.py:153, 178, 181).  Only numpy and scipy are imported here: the
functions run in worker PROCESSES (hostops.pvalue_pool) -- scipy's Boost-backed ufuncs keep the interpreter lock
(four threads on `_hypergeom_sf` take four times as long as one), so threads cannot share this work, and
hypergeom.sf at 20-150 us per loop is what bounds the byte-identical mode.
"""
import numpy as np
from scipy.stats import binom, hypergeom, poisson


def floor300(p):
    """max([1e-300, p]) as the reference wraps every tail probability."""
    return np.maximum(1e-300, p)


def anchor_sig(count, wide, length, rpb):
    """estAnchorSig (callCisLoops.py:226-247) from the two counts: returns (p, es)."""
    r = (wide - count) / 5 / 2
    c = np.maximum(np.maximum(r, rpb * length), 1.0)
    p = floor300(poisson.sf(count - 1.0, c))
    return p, count / c


def tails(rab, ra, rb, Nv, rv, mrabs, mbps, la, lb, a0, a1, a2, a3, cis=True):
    """float64[n, 7]: hypergeometric, Poisson and binomial tails of the loops, Poisson tail and enrichment of either
    anchor.  Every argument is an array of n rows (Nv / rv: the file's PET count and PETs per bp, per row)."""
    hyp = hypergeom.sf(rab - 1.0, Nv, ra, rb)
    pop = poisson.sf(rab - 1.0, mrabs)
    nbp = binom.sf(rab - 1.0, ra * rb - rab if cis else ra * rb, mbps)
    px, esx = anchor_sig(a0, a1, la, rv)
    py, esy = anchor_sig(a2, a3, lb, rv)
    return np.stack([hyp, pop, nbp, px, esx, py, esy], axis=1)


# ---------------------------------------------------------------------------------------------- worker process
# `python -m cloops2_b200.pvalues --worker RFD WFD`: length-prefixed pickles on the inherited pipe RFD ((args tuple,
# cis) per task), the float64[n, 7] result of tails() on the pipe WFD, until RFD closes.  The pipes are the worker's
# own (not stdin / stdout, which anything in the interpreter may write to).  Started by hostops.pvalue_pool(); a plain
# child process (fresh interpreter, numpy + scipy only), so nothing of the parent -- its CUDA context, its threads, its
# __main__ module -- is inherited or re-imported.
def _read_exact(f, n):
    buf = b""
    while len(buf) < n:
        part = f.read(n - len(buf))
        if not part:
            return None
        buf += part
    return buf


def send_msg(f, obj):
    import pickle
    import struct
    data = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    f.write(struct.pack("<Q", len(data)))
    f.write(data)
    f.flush()


def recv_msg(f):
    import pickle
    import struct
    head = _read_exact(f, 8)
    if head is None:
        return None
    data = _read_exact(f, struct.unpack("<Q", head)[0])
    return None if data is None else pickle.loads(data)


def _worker_main(rfd, wfd):
    import os
    fin, fout = os.fdopen(rfd, "rb"), os.fdopen(wfd, "wb")
    while True:
        task = recv_msg(fin)
        if task is None:
            return
        args, cis = task
        try:
            send_msg(fout, ("ok", tails(*args, cis=cis)))
        except Exception as e:               # reported to the parent, which evaluates the chunk itself
            send_msg(fout, ("err", repr(e)))


if __name__ == "__main__":
    import sys
    if len(sys.argv) == 4 and sys.argv[1] == "--worker":
        _worker_main(int(sys.argv[2]), int(sys.argv[3]))
