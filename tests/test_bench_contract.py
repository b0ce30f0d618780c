"""CPU-side checks of the measurement contract and of the oracle quarantine (no GPU needed)."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` times the oracle port on the host and prints ONE JSON line with the contract keys."""
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--workload", "c3", "--scale", "0.2"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["dtype"] == "f64" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] > 0 and cb["sample"]
    assert set(cb["buckets_s"]) == {"grad", "limiter", "flux", "rk"}          # the reference's four timers (src/mainparam.f90:15-16)
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "not the reference algorithm" in cb["all_cores_variant"].get("note", "") or "unavailable" in cb["all_cores_variant"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load oracle/ (tier rule 3)."""
    pat = re.compile(r"\boracle\b")
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "fvs2d_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(base, f), errors="ignore").read()
                for i, line in enumerate(txt.splitlines(), 1):
                    if pat.search(line) and not re.search(r"(import|include|dlopen|CDLL|load)", line) is None:
                        bad.append(f"{f}:{i}: {line.strip()}")
    assert not bad, bad
    hdr = open(os.path.join(ROOT, "include", "fvs2d_gpu.h")).read()
    assert 'extern "C"' in hdr and not re.search(r"#include\s*<(torch|ATen|c10)", hdr)   # plain C types only
    # the graft entry uses the oracle only inside smoke()
    src = open(os.path.join(ROOT, "__graft_entry__.py")).read()
    head = src.split("def smoke")[0]
    assert not re.search(r"^\s*(from|import)\s+oracle", head, re.M)
