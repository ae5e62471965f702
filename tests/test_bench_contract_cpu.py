"""bench.py's reference arm and JSON contract on the CPU (no GPU needed): the line the driver
parses must carry the contract keys, and the CPU arm must agree with the oracle port (it
asserts that internally while timing the reference's compiled CEUpdater)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--gpus", "1", "--steps", "1", "--warmup", "0", "--ref-moves", "3000"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mc_trial_moves_per_sec"
    assert d["unit"] == "moves/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "moves/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert "BASELINE configs[1]" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: without a CUDA device the product path raises."""
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1",
                          "--warmup", "0", "--no-cpu-baseline", "--no-extra"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0
    assert out.stdout.strip() == ""


def test_profiler_figures_only_from_this_build(tmp_path, monkeypatch):
    """roofline.traffic / roofline_issue come from profiles/traffic.json only when it was captured on
    this build (kernel source hash); a stale file must yield null with the reason, never canned
    counters, and the issue-slot roofline must follow from its inputs."""
    sys.path.insert(0, ROOT)
    import bench
    from cemc_b200 import _lib
    prof = tmp_path / "profiles"
    prof.mkdir()
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    clocks = {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 3}
    # stale capture
    (prof / "traffic.json").write_text(json.dumps({"kernel_source_sha": "0" * 16,
                                                    "C2": {"dram_bytes_per_launch": 1, "warp_instructions_per_move": 300.0}}))
    roof, issue = bench.rooflines("C2", 258, 5120000, 3.2e-3, 1.6e9, clocks)
    assert roof["traffic"] is None and issue is None and "another build" in roof["traffic_source"]
    assert abs(roof["achieved"] - 258 * 5120000 / 3.2e-3 / 1e9) < 1e-9 and roof["frac"] == roof["achieved"] / roof["peak"]
    # capture of this build
    (prof / "traffic.json").write_text(json.dumps({"kernel_source_sha": _lib.source_hash(),
                                                    "C2": {"dram_bytes_per_launch": 453888, "warp_instructions_per_move": 289.2,
                                                           "issue_slots_busy_pct": 51.0, "source": "x"}}))
    roof, issue = bench.rooflines("C2", 258, 5120000, 3.2e-3, 1.6e9, clocks)
    assert roof["traffic"] == 453888
    assert issue["bound"] == "issue" and issue["peak"] == 148 * 4 * 1965.0e6
    assert abs(issue["frac"] - 289.2 * 1.6e9 / (148 * 4 * 1965.0e6)) < 1e-12
