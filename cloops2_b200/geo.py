"""
Loop geometry of the callLoops path (cLoops2/geo.py:65-112), object-level for API compatibility.  The array-level
equivalents used by the fast path live in cloops2_b200/hostops.py.
"""


def checkAnchorOverlap(xa, xb, ya, yb):
    """Closed-interval overlap on one chromosome (geo.py:65-73)."""
    return (ya <= xa <= yb) or (ya <= xb <= yb) or (xa <= ya <= xb) or (xa <= yb <= xb)


def checkLoopOverlap(loopa, loopb):
    """Both anchors overlap and chromosomes match (geo.py:76-87)."""
    if loopa.chromX != loopb.chromX or loopa.chromY != loopb.chromY:
        return False
    return (checkAnchorOverlap(loopa.x_start, loopa.x_end, loopb.x_start, loopb.x_end)
            and checkAnchorOverlap(loopa.y_start, loopa.y_end, loopb.y_start, loopb.y_end))


def _coords(loop):
    return (loop.chromX, loop.x_start, loop.x_end, loop.chromY, loop.y_start, loop.y_end)


def combineLoops(loops, loops_2):
    """Merge candidate dicts of two clustering runs (geo.py:90-112): later runs only add unseen coordinates."""
    for key, cand in loops_2.items():
        if key not in loops:
            loops[key] = cand
            continue
        seen = {_coords(lp) for lp in loops[key]}
        loops[key].extend(lp for lp in cand if _coords(lp) not in seen)
    return loops
