"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`, the oracle port of the reference's
PyTorch-CPU path timed on the host cores) prints ONE JSON line with the keys the driver reads, ranks other than 0 exit
quietly, and the roofline traffic capture under profiles/ is the one taken from the current kernel sources."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "0"], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                          text=True, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "UNet-64 train images/sec" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith("{")]


def test_committed_traffic_capture_matches_the_kernel_sources():
    """`roofline.traffic` comes from profiles/r02_tc_traffic.json, which is only reported while the tcgen05 kernel sources
    are the ones it was captured from; a commit that changes a kernel without re-capturing would turn it into null."""
    sys.path.insert(0, ROOT)
    import bench
    traffic, note = bench.conv_traffic()
    assert traffic is not None and traffic > 0, note
