"""CPU: the parts of bench.py that run without a GPU — the reference arm's JSON line (the driver parses it), workload
table consistency with BASELINE.json / SURVEY §8d, the ncu-traffic lookup, and the clocks parser."""
import json
import re
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "segformer_b2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "tiles/s" and d["higher_is_better"] is True
    assert d["metric"] == "512x512 multi-band tiles/sec (train fwd+bwd)" and d["value"] > 0
    assert d["config"]["workload"] == "segformer_b2_3band_512_k5_b16"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1


def test_reference_arm_other_ranks_do_no_work():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_workloads_match_the_baseline_configs():
    import bench
    base = json.loads((ROOT / "BASELINE.json").read_text())
    assert "tiles/sec" in base["metric"]
    w = bench.WORKLOADS
    assert bench.WORKLOAD is w["unetpp_r50"]  # the default = configs[1], the single-GPU configuration of the metric
    assert (w["unetpp_r50"]["bands"], w["unetpp_r50"]["tile"], w["unetpp_r50"]["batch_per_gpu"]) == (4, 512, 32)
    assert (w["segformer_b2"]["bands"], w["segformer_b2"]["batch_per_gpu"]) == (3, 16)
    assert w["dofa_base"]["bands"] == 6 and len(w["dofa_base"]["wavelengths"]) == 6
    assert w["segformer_b5_infer"]["raster"] == 10000 and w["segformer_b5_infer"]["family"] == "infer"
    # training = 3 x forward GFLOP per tile
    assert abs(w["unetpp_r50"]["train_gflop_per_tile"] - 3 * 460.24) < 0.5
    assert abs(w["segformer_b2"]["train_gflop_per_tile"] - 3 * 121.03) < 0.5
    assert abs(w["dofa_base_unfrozen"]["train_gflop_per_tile"] - 3 * 666.10) < 0.5


def test_dram_traffic_lookup_uses_the_newest_committed_pass():
    import bench
    for fam in ("unetpp", "segformer", "dofa"):
        per_launch, src = bench._dram_traffic(fam)
        assert per_launch and per_launch > 1e6 and re.match(r"profiles/r0\d_run\d+_dram_traffic_", src) and (ROOT / src).exists()
    assert bench._dram_traffic("unetpp")[1].startswith("profiles/r02_")  # the newest round wins
    assert bench._dram_traffic("no_such_family") == (None, None)


def test_clock_sampler_parses_nvidia_smi_rows():
    import bench
    c = bench.ClockSampler(0)
    assert c.stop()["reasons"] == ["nvidia-smi unavailable"]  # never started

    class _P:
        def terminate(self):
            pass
    c.proc = _P()
    c.rows = [["0", "1965", "1965", "850.1", "0x0", "Not Active", "Not Active", "Not Active", "Not Active"],
              ["0", "1890", "1965", "990.0", "0x4", "Not Active", "Not Active", "Not Active", "Active"],
              ["0", "1965", "1965", "900.0", "0x0", "Not Active", "Not Active", "Not Active", "Not Active"]]
    out = c.stop()
    assert out == {"sm_mhz": 1965, "sm_max_mhz": 1965, "reasons": ["sw_power_cap"], "samples": 3}
