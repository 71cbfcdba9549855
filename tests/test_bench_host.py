"""Host-side logic of bench.py that needs no GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_only_the_json_line_reaches_stdout():
    """Native libraries print to file descriptor 1 (NCCL's version banner did, in front of the JSON line); bench.py keeps
    the real stdout for the one line the driver parses and sends everything else to stderr."""
    code = (
        "import os, sys\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import bench\n"
        "bench.claim_stdout()\n"
        "os.write(1, b'native noise\\n')\n"
        "print('python noise')\n"
        "bench.emit('{\"ok\": 1}')\n"
    )
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"ok": 1}\n'
    assert "native noise" in r.stderr and "python noise" in r.stderr


def test_reference_arm_prints_one_json_line():
    sys.path.insert(0, ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2_n10",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    from oracle.ref_loader import reference_available
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == ("reference" if reference_available() else "port")
    assert d["steps"] == 1 and d["warmup"] == 1                      # the arm honours --steps / --warmup
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["metric"].startswith("agent-steps/sec") and d["higher_is_better"] is True


def test_algorithmic_bytes_match_the_survey():
    sys.path.insert(0, ROOT)
    import bench
    w = bench.WORKLOADS["c4_n1000"]
    # forward 4N^2 + 4GN + 4CN per instance = 6.56 MB at N = 1000 -> 3.359 GB per batch of 512
    assert bench.alg_bytes(w, "fwd") == 512 * (4 * 1000 * 1000 + 4 * 128 * 1000 + 4 * 512 * 1000) == 3358720000
    assert bench.alg_bytes(w, "train") - bench.alg_bytes(w, "fwd") == 512 * 3072 * 1000
