"""The bench.py JSON lines committed under profiles/ (measured on B200 through gpurun) carry every key of the bench
contract, and the CPU arm (`bench.py --impl reference`), which needs no GPU, prints the same shape here."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config"]


def last_line(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


@pytest.mark.parametrize("name,n", [("r01e_bench_n1.json", 1), ("r01e_bench_n2.json", 2), ("r01d_bench_n4.json", 4), ("r01d_bench_n8.json", 8),
                                    ("r02b_bench_n1.json", 1), ("r02b_bench_n2.json", 2), ("r02b_bench_n4.json", 4), ("r02b_bench_n8.json", 8),
                                    ("r02c_bench_n1.json", 1), ("r02c_bench_n2.json", 2), ("r02c_bench_n8.json", 8)])
def test_committed_gpu_records(name, n):
    d = last_line(os.path.join(REPO, "profiles", name))
    for k in BASE + ["clocks", "gpu_launches", "e2e", "roofline", "cpu_baseline"]:
        assert k in d, k
    assert d["n_gpus"] == n and d["unit"] == "Mobs/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["value"] > 0
    assert abs(d["value"] - d["config"]["observations"] * d["steps"] / (d["ms_per_step"] * d["steps"] * 1e-3) / 1e6) < 1e-6 * d["value"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"]
    assert 0 < d["e2e"]["value"] < d["value"] and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert d["roofline_hbm"]["bound"] == "hbm" and d["roofline_hbm"]["frac"] >= 0.6      # the north star's >= 60 % of HBM
    assert set(d["clocks"]["reasons"]).isdisjoint({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"})
    if n == 1:
        c = d["cpu_baseline"]
        for k in ("value", "unit", "cores", "kind", "sample"):
            assert k in c
        assert c["kind"] in ("port", "reference") and c["cores"] >= 1
    else:
        assert d["cpu_baseline"] is None          # rank 0 at N = 1 only
    if name.startswith("r02"):                    # round 2: every line carries a parity object and says which data plane ran
        assert d["parity"]["pass"] is True and all(st["same_accept_reject_sequence_and_termination"] for st in d["parity"]["stages"])
        assert d["schur_reduce"] == ("single" if n == 1 else "p2p") and d["launch_mode"].startswith("CUDA graphs")


def test_reference_arm_runs_on_the_cpu():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--frames-per-gpu", "12"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    for k in BASE + ["impl", "e2e", "cpu_baseline"]:
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Mobs/s" and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["single_thread_value"] > 0
    assert d["config"]["frames"] == 12 and d["config"]["observations"] == d["cpu_baseline"]["observations_timed"]   # same config, same shard


def test_committed_launch_list_holds_only_this_librarys_kernels():
    """profiles/r02c_ncu_launches_bench_steps2.csv (ncu launch list of the bench command): every kernel that ran is one
    of libmcba's own - no library kernel (cuBLAS, cuSOLVER, NCCL, torch) anywhere on the path - and the three
    heavyweights of an LM iteration are there."""
    import csv
    import re
    rows = list(csv.reader(open(os.path.join(REPO, "profiles", "r02c_ncu_launches_bench_steps2.csv"))))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    kn = rows[hi].index("Kernel Name")
    names = {re.sub(r"\(.*", "", r[kn]).replace("void ", "").strip() for r in rows[hi + 1:] if len(r) > kn}
    src = "".join(open(os.path.join(REPO, "mc-calib_b200", "csrc", f)).read()
                  for f in os.listdir(os.path.join(REPO, "mc-calib_b200", "csrc")) if f.endswith((".cu", ".cuh")))
    ours = set(re.findall(r"\b(k_[a-z0-9_]+)\s*\(", src))
    assert len(names) >= 15
    for n in names:
        base = re.sub(r"<.*", "", n).replace("mcba::", "")
        assert base in ours, n
    assert {"k_syrk2", "k_chol_tiles", "k_observations_pair"} <= {re.sub(r"<.*", "", n).replace("mcba::", "") for n in names}
