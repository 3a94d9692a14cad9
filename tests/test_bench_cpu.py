"""bench.py contract pieces that need no GPU: the reference arm (the unmodified reference on this box's cores, or
"unavailable" where oracle/_ref was not built) prints ONE JSON line with the agreed keys; non-zero ranks stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "1",
                           "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_line():
    r = run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    if "unavailable" in d:
        return
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config",
                "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["metric"] == "kmers_counted_per_sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0


def test_reference_arm_other_ranks_are_silent():
    r = run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_both_arms_describe_the_same_config():
    """`config` is what the workload IS: the reference arm prints the same dictionary our arm computes (the driver compares
    them), whatever the number of ranks."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    r = run()
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.strip()][0])
    if "unavailable" in d:
        return
    assert d["config"] == bench.workload_config("tiny", 1)
    c8 = bench.workload_config("c2_150Mbp_10kbp", 8)
    assert c8["genome_len_total"] == 40_000_000 and c8["reads_per_gpu"] == 15_000 and c8["kmers_per_gpu"] == 149_550_000
    assert bench.workload_config("c3_30Gbp", 8)["scaling"] == "strong" and bench.workload_config("c3_30Gbp", 8)["reads_per_gpu"] == 375_000
