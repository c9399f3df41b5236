"""The product package must never import or call the oracle (or any CPU fallback)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_does_not_reference_oracle():
    pkg = os.path.join(ROOT, "cloops2_b200")
    bad = []
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M) or "cl2_oracle" in txt or "orc_" in txt:
                    bad.append(f)
    assert bad == []
