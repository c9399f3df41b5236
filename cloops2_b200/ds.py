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
