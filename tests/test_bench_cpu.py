"""Host-side contract of bench.py that needs no GPU: the reference arm's JSON line, the loud failure of the product arm
without a CUDA device, and the child run that adds configs[2] at the same N to every default line."""
import json
import os
import subprocess
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(*argv, env=None, timeout=600):
    return subprocess.run([sys.executable, BENCH, *argv], capture_output=True, text=True, timeout=timeout,
                          env=env, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train scans/sec" and d["unit"] == "scans/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("configs[1]")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", env=env, timeout=120)
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return
    r = _run("--no-scaling-baseline", "--steps", "1", timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_configs2_child_run(monkeypatch):
    sys.path.insert(0, ROOT)
    import bench
    seen = {}
    child_line = {"value": 1234.5, "unit": "scans/s", "ms_per_step": 12.9, "n_gpus": 2,
                  "e2e": {"value": 1200.0, "unit": "scans/s", "h2d_bytes_per_step": 1, "d2h_bytes_per_step": 4},
                  "config": {"workload": "configs[2]: kitti", "scans_per_gpu": 8, "parallelism": "dp2",
                             "cuda_graph": "whole step"}, "clocks": {"sm_mhz": 1965.0}}

    def fake_run(cmd, env=None, **kw):
        seen["cmd"], seen["env"] = cmd, env
        return types.SimpleNamespace(stdout="noise\n" + json.dumps(child_line) + "\n", stderr="", returncode=0)

    monkeypatch.setattr(bench.subprocess, "run", fake_run)
    monkeypatch.setenv("TORCHELASTIC_USE_AGENT_STORE", "True")
    monkeypatch.setenv("MASTER_PORT", "29611")
    args = types.SimpleNamespace(gpus=2, steps=20, warmup=5, fusion="middle", fmap_format="channels_last")
    out = bench.run_configs2_child(args, rank=0, world=2)
    assert out["value"] == 1234.5 and out["n_gpus"] == 2 and out["workload"].startswith("configs[2]")
    cmd, env = seen["cmd"], seen["env"]
    assert cmd[cmd.index("--workload") + 1] == "kitti" and cmd[cmd.index("--gpus") + 1] == "2"
    assert "--no-roofline" in cmd and "--no-cpu-baseline" in cmd
    # the children rendezvous among themselves: own port, own store (not torchrun's agent store)
    assert env["MASTER_PORT"] == "29612" and "TORCHELASTIC_USE_AGENT_STORE" not in env
    assert bench.run_configs2_child(args, rank=1, world=2) is None           # only rank 0 reports
    # a failing child is reported in the key, never raised
    monkeypatch.setattr(bench.subprocess, "run",
                        lambda *a, **k: types.SimpleNamespace(stdout="", stderr="boom", returncode=3))
    assert "error" in bench.run_configs2_child(args, rank=0, world=2)
    # one GPU: the environment is passed through untouched
    args.gpus = 1
    monkeypatch.setattr(bench.subprocess, "run", fake_run)
    bench.run_configs2_child(args, rank=0, world=1)
    assert seen["env"]["MASTER_PORT"] == "29611"
