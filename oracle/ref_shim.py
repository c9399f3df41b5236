"""
TEST INFRASTRUCTURE ONLY -- import shim for the live Python reference at /root/reference.

The reference (cLoops2 0.0.5) is pure Python and importable in the build container once its
plotting imports (cLoops2/settings.py:16-26, pulled in by callCisLoops.py:45) are stubbed.
It is used ONLY to (i) validate the C restatement in oracle/cl2_oracle.c and (ii) generate the
golden vectors committed under tests/golden/ (oracle/gen_golden.py).  /root/reference does not
exist on the GPU box; nothing in tests marked `gpu`, smoke() or bench.py imports this module.
"""
import os
import sys
import types
import logging
from unittest import mock

REF_ROOT = os.environ.get("CL2_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "cLoops2"))


_loaded = {}


def load():
    """Import the reference hot-path modules; returns a namespace of the callables we pin against."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REF_ROOT)
    for name in ("matplotlib", "matplotlib.colors", "matplotlib.pyplot", "matplotlib.ticker",
                 "matplotlib.patches", "matplotlib.gridspec", "matplotlib.backends",
                 "matplotlib.backends.backend_pdf", "seaborn", "pylab", "pyBigWig", "mpl_toolkits",
                 "mpl_toolkits.axes_grid1", "mpl_toolkits.axes_grid1.inset_locator"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = mock.MagicMock()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from cLoops2.blockDBSCAN import blockDBSCAN
    from cLoops2.cDBSCAN2 import cDBSCAN
    from cLoops2 import callCisLoops as cis
    from cLoops2 import callTransLoops as trans
    from cLoops2 import ds, geo, io
    log = logging.getLogger("cl2_ref")
    log.addHandler(logging.NullHandler())
    log.propagate = False
    cis.logger = log
    ns = types.SimpleNamespace(blockDBSCAN=blockDBSCAN, cDBSCAN=cDBSCAN, cis=cis, trans=trans,
                               ds=ds, geo=geo, io=io, logger=log)
    _loaded["ns"] = ns
    return ns
