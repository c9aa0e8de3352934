"""CPU: the JSON lines bench.py printed on the B200 (committed under profiles/) carry every key of the bench contract.

The bench itself needs a GPU; this guards the contract (and the committed evidence) against drifting apart: the
newest `r*_bench_n1.json` / `r*_bench_reference_arm.json` / multi-GPU lines are parsed and checked key by key."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _newest(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    if not files:
        pytest.skip("no %s under profiles/" % pattern)
    with open(files[-1]) as f:
        return json.loads(f.read().strip().splitlines()[-1]), os.path.basename(files[-1])


BASE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches")


def test_single_gpu_line():
    d, name = _newest("r*_bench_n1.json")
    for k in BASE_KEYS + ("clocks", "roofline", "cpu_baseline"):
        assert k in d, (name, k)
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert d["metric"].split(" (")[0] in base["metric"] and d["unit"] == "evals/s"
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert "workload" in d["config"] and "configs[1]" in d["config"]["workload"] and "l2" in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] == d["steps"]
    assert abs(d["value"] - 1e6 * 1e3 / d["ms_per_step"]) / d["value"] < 1e-6
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 40_000_000 and e["d2h_bytes_per_step"] == 8_000_000 and 0 < e["value"] < d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - 944.0 * d["value"] / 1e9) / r["achieved"] < 1e-6 and r["traffic"] > 0
    assert 0.5 < r["l1_gather_bound"]["frac"] < 1.0
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["unit"] == "evals/s" and c["value"] > 0 and c["sample"]
    clk = d["clocks"]
    assert clk["sm_mhz"] and clk["sm_max_mhz"] and not [x for x in clk["reasons"] if "slowdown" in x]
    assert d["value"] >= 1e8                                   # BASELINE.json north_star target


def test_reference_arm_line():
    d, name = _newest("r*_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["value"] > 0, name
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


@pytest.mark.parametrize("n", [2, 4, 8])
def test_multi_gpu_lines(n):
    d, name = _newest("r*_bench_n%d.json" % n)
    for k in BASE_KEYS:
        assert k in d, (name, k)
    assert d["n_gpus"] == n and d["scaling"] == "weak"
    assert abs(d["value"] - n * 1e6 * 1e3 / d["ms_per_step"]) / d["value"] < 1e-6
    g = d["allgather"]
    assert g["collective"].startswith("ncclAllGather") and g["fused_peer_store"]["identical_to_nccl"] is True
    assert g["fused_peer_store"]["ms_per_step"] < g["ms_per_step_with_gather"]
    for k in ("binary_1e6_rows_sharded", "catalog_10k_stars_sharded"):
        assert d["alt"][k]["fused_identical_to_nccl"] is True
