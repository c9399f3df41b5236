"""
Drop-in for cLoops2/cDBSCAN2.py: `cDBSCAN(mat, eps, minPts).labels`, used when `-hic` is given
(callCisLoops.py:499-502).  Same protocol as blockDBSCAN.
"""
import numpy as np

from .blockDBSCAN import blockDBSCAN


class cDBSCAN(blockDBSCAN):
    @staticmethod
    def _run(pts, eps, minPts):
        if pts.n == 0:
            return np.zeros(0, dtype=np.int32), 0
        return pts.cdbscan2(eps, minPts)
