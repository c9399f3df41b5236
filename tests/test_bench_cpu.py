"""bench.py on a CPU-only host: workload definition, the parity comparison of loop rows, and the time-bounded
reference arm (`--impl reference` runs the oracle port -- test infrastructure -- on the host cores)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_workload_definitions():
    c2 = bench.workload_jobs(bench.CONFIGS["c2"], 100e6, 1, "weak")
    assert len(c2) == 24 and abs(sum(bench.job_pets(j) for j in c2) - 100e6) < 100
    assert bench.job_pets(c2[0]) == max(bench.job_pets(j) for j in c2)               # chr1 is the largest file
    assert len(bench.workload_jobs(bench.CONFIGS["c2"], 100e6, 8, "weak")) == 192        # N genomes on N GPUs
    assert len(bench.workload_jobs(bench.CONFIGS["c2"], 100e6, 8, "strong")) == 24       # one genome over N GPUs
    assert len(set(j[1] for j in bench.workload_jobs(bench.CONFIGS["c2"], 100e6, 8, "weak"))) == 192
    c4 = bench.workload_jobs(bench.CONFIGS["c4"], 200e6, 1, "weak")
    assert len(c4) == 276 and all(j[0] == "trans" for j in c4)
    c3 = bench.workload_jobs(bench.CONFIGS["c3"], 1e9, 8, "weak")
    assert len(c3) == 24                                                                 # one genome, cut into shards by the caller
    small = bench.cpu_sample_files(bench.CONFIGS["c2"], 5.5e6, c2)
    assert sum(bench.job_pets(j) for j in small) <= 5.5e6 and len(small) == 3


def test_compare_rows_text_identity_and_tail_tolerance():
    row = ["loop_chr1-chr1-0", "chr1", "100", "200"] + ["x"] * 14 + ["1e-12", "2e-30", "3e-8", "4e-9", "5e-7", "1"]
    assert len(row) == 24
    same = [list(row)]
    assert bench.compare_rows(same, [row]) == (True, 0.0)
    near = [list(row)]
    near[0][19] = repr(2e-30 * (1 + 1e-9))
    ok, worst = bench.compare_rows(near, [row])
    assert ok and 0 < worst < 2e-9
    far = [list(row)]
    far[0][19] = repr(2e-30 * (1 + 1e-3))
    assert bench.compare_rows(far, [row])[0] is False
    other = [list(row)]
    other[0][2] = "101"                                  # any other column must be identical text
    assert bench.compare_rows(other, [row])[0] is False
    assert bench.compare_rows([], [row])[0] is False


def test_reference_arm_is_time_bounded_and_prints_the_contract_line():
    t = time.time()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--ref-budget-s", "6"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    dt = time.time() - t
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.strip()]
    assert len(lines) == 1                               # exactly one JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "PETs/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "PETs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["steps"] >= 1 and d["warmup"] == 1 and d["value"] > 0
    assert d["config"]["workload"].startswith("BASELINE configs[1]")
    assert dt < 240, "reference arm took %.0f s for a 6 s budget" % dt


def test_recorded_bench_line_carries_the_contract_keys():
    """The last default bench line recorded on a B200 (profiles/) has every key the measurement contract names."""
    d = json.loads(open(os.path.join(ROOT, "profiles", "r2e_bench_1gpu_c2_final.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == bench.METRIC and d["unit"] == "PETs/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("BASELINE configs[1]") and d["data"] == "synthetic" and d["dtype"] == "int32"
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 16 * 99e6 and d["e2e"]["d2h_bytes_per_step"] > 0      # copies inside the timed region
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"]) and d["roofline"]["bound"] == "hbm"
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-12
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] == "port"
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"]) and d["gpu_launches"] > 0
    assert d["parity_check"]["status"] == "equal"
    assert abs(d["value"] - d["job"]["pets"] / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-9
