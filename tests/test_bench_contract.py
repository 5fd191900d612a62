"""bench.py's driver-facing contract, as far as it can be checked without a GPU: the reference arm's JSON line, the
silent non-zero ranks of a torchrun launch, and the parsing of the nvidia-smi clock samples."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                          text=True, timeout=900, env=e)


def test_reference_arm_prints_one_json_line_on_the_cuda_arms_config():
    res = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] and d["unit"] == "Gelem/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["vs_baseline"] is None and d["data"] == "synthetic"
    cfg = d["config"]
    assert cfg["workload"] == "kodak_sweep" and cfg["n_per_unit"] == 49152 and cfg["units_per_step"] == 1010
    assert cfg["elements_per_step"] == 1010 * 49152           # the whole step, not a sample of it
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "1010 of the 1010 units" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Gelem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert abs(d["value"] - cfg["elements_per_step"] / (d["ms_per_step"] * 1e-3) / 1e9) < 1e-9 * max(1.0, d["value"])


def test_reference_arm_other_ranks_exit_quietly():
    res = _run(["--impl", "reference", "--gpus", "2"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_clock_sampler_parses_and_falls_back_per_gpu():
    import bench

    s = bench.ClockSampler(0, 2)
    s.proc = type("P", (), {"terminate": lambda self: None})()
    row = "{}, {}, 1965, 600.0, Not Active, Not Active, Not Active, {}"
    s.rows = [(9.0, row.format(0, 1900, "Not Active")), (9.0, row.format(1, 1800, "Not Active")),   # before the region
              (10.5, row.format(0, 1950, "Active")),                                                  # inside: GPU 0 only
              (12.0, row.format(1, 1700, "Not Active")), (12.0, "garbage")]                           # after / unparsable
    out = s.stop(10.0, 11.0)
    assert out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]
    assert out["sm_mhz_by_gpu"] == [1950.0, 1750.0]         # GPU 1 had no sample inside: median of its other samples
    assert out["sm_mhz"] == 1750.0 and out["samples_in_timed_region"] == 1
