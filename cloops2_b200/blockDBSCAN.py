"""
Drop-in for cLoops2/blockDBSCAN.py: `blockDBSCAN(mat, eps, minPts).labels` (dict rowId -> clusterId, clustered
points only, reference cluster numbering), computed by the sm_100a kernels of libcl2.so.
Call sites in the reference: callCisLoops.py:87-88, callTransLoops.py:53-54, callPeaks.py:97-98.
"""
import numpy as np

from . import _lib


class blockDBSCAN:
    def __init__(self, mat, eps, minPts):
        """mat: [pointId, X, Y] rows (int or integral float, as the reference builds at callCisLoops.py:67-72)."""
        self.eps = eps
        self.minPts = minPts
        mat = np.asarray(mat)
        ids = mat[:, 0]
        xy = np.ascontiguousarray(mat[:, 1:3].astype(np.int64))
        ctx = _lib.default_context()
        pts = _lib.Points(ctx, xy)
        try:
            lab, self.n_clusters = self._run(pts, int(eps), int(minPts))
        finally:
            pts.close()
        sel = np.nonzero(lab >= 0)[0]
        self.labels = dict(zip(ids[sel].tolist(), lab[sel].tolist()))

    @staticmethod
    def _run(pts, eps, minPts):
        if pts.n == 0:
            return np.zeros(0, dtype=np.int32), 0
        return pts.block_dbscan(eps, minPts)
