"""
Data records of the callLoops path.  `Loop` mirrors cLoops2/ds.py:245-281 slot for slot so objects produced here can
be handed to the reference's own writers / downstream commands unchanged.
"""


class Loop(object):
    """Used to store loop related information (same 26 slots as the reference's Loop)."""
    __slots__ = [
        "id", "chromX", "chromY", "x_start", "x_end", "x_center", "ra", "y_start", "y_end", "y_center", "rb", "rab",
        "cis", "distance", "density", "ES", "P2LL", "FDR", "hypergeometric_p_value", "poisson_p_value",
        "binomial_p_value", "x_peak_poisson_p_value", "x_peak_es", "y_peak_poisson_p_value", "y_peak_es",
        "significant",
    ]

    def __str__(self):
        return str(self.__class__) + ": <%s:%s-%s><%s:%s-%s>" % (self.chromX, self.x_start, self.x_end, self.chromY,
                                                                 self.y_start, self.y_end)


class DiffLoop(object):
    """Differentially enriched loop record (same 31 slots as the reference's DiffLoop, cLoops2/ds.py:285-326)."""
    __slots__ = [
        "id", "chromX", "chromY", "x_start", "x_end", "x_center", "y_start", "y_end", "y_center", "distance", "size",
        "raw_trt_ra", "raw_trt_rb", "raw_con_ra", "raw_con_rb", "scaled_trt_ra", "scaled_trt_rb", "raw_trt_rab",
        "raw_con_rab", "raw_trt_mrab", "raw_con_mrab", "scaled_trt_rab", "scaled_trt_mrab", "trt_density",
        "con_density", "trt_es", "con_es", "raw_fc", "scaled_fc", "poisson_p_value", "significant",
    ]

    def __str__(self):
        return str(self.__class__) + ": <%s:%s-%s><%s:%s-%s>" % (self.chromX, self.x_start, self.x_end, self.chromY,
                                                                 self.y_start, self.y_end)


class Peak(object):
    """Peak record (same 20 slots as the reference's Peak, cLoops2/ds.py:212-241)."""
    __slots__ = [
        "id", "chrom", "start", "end", "summit", "counts", "length", "density", "poisson_p_value", "enrichment_score",
        "control_counts", "control_local_counts", "control_density", "control_scaled_counts",
        "poisson_p_value_vs_control", "enrichment_score_vs_control", "up_down_counts", "control_up_down_counts",
        "p_value_mean", "significant",
    ]

    def __str__(self):
        return str(self.__class__) + ": <%s:%s-%s>" % (self.chrom, self.start, self.end)
