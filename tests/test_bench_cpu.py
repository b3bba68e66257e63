"""CPU checks of bench.py's contract: the reference arm prints one JSON line with the agreed keys; the GPU arm
refuses to run without a CUDA device (no CPU fallback of the product path)."""
import json
import os
import subprocess
import sys

from helpers import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["metric"] == base["metric"] and d["unit"] == "tokens/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert "workload" in d["config"] and "model" not in d["config"]
    # the reference arm runs on the GPU arm's own config: the same dict, and its step includes both ops of the path
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.bench_config(1)
    assert "reshape_and_cache" in cb["sample"] and "softmax" in cb["sample"]


def test_kernel_cells_cover_baseline_configs_1_to_3():
    sys.path.insert(0, ROOT)
    import bench
    names = [n for n, _ in bench.KERNEL_CELLS]
    assert len(names) == 3 + 9 and names[0].startswith("metric")
    assert sum(n.startswith("configs[3]") for n in names) == 9
    shapes = dict(bench.KERNEL_CELLS)
    assert shapes["configs[1] 64x512 12x64 blk16"] == (64, 12, 64, 16, 512)
    assert shapes["configs[3] 256x8k 20x64 blk8"] == (256, 20, 64, 8, 8192)


def test_reference_arm_non_zero_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, BENCH, "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                         timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in out.stderr
