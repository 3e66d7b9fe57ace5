"""bench.py contract checks that need no GPU: the reference arm (CPU oracle port on the host cores) prints one JSON line
with the keys the driver reads, and the GPU arm's helpers are importable without CUDA."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_step_volumes_per_s_128cubed" and d["unit"] == "volumes/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_window_enumeration_of_config5():
    """custom_callback.py:127-162 semantics on BASELINE config 5: 256 calls (147 unique) without padding, 864 (605) with
    complete=True, padFactor=0.25."""
    sys.path.insert(0, ROOT)
    from van_gan_b200.custom_callback import window_starts
    def count(shape):
        st = [(a, b, c) for a in window_starts(shape[0], 128, 64) for b in window_starts(shape[1], 128, 64)
              for c in window_starts(shape[2], 128, 64)]
        return len(st), len(set(st))
    assert count((512, 512, 256)) == (256, 147)
    assert count((768, 768, 384)) == (864, 605)
