"""CPU: the `--impl reference` arm of bench.py (the unmodified reference on the host cores when its copy is staged
under oracle/_ref/py or /root/reference exists, else the oracle port) prints ONE JSON line with the keys the driver
reads; the flop / byte model of the roofline is importable and consistent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-sample", "8", "--model", "pendulum"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "solves/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("AL-MPC solves/sec") and d["value"] > 0 and d["ms_per_step"] > 0
    for k in ("n_gpus", "steps", "warmup", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert "workload" in d["config"] and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    sys.path.insert(0, ROOT)
    from oracle import ref_import

    assert cb["kind"] == ("reference" if ref_import.available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["scaling"] == "strong" and d["config"]["global_batch"] == 65536
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_flop_and_byte_model():
    sys.path.insert(0, ROOT)
    import bench

    fl = bench.algorithmic_flops("rex_quadrotor", 5)
    by = bench.algorithmic_bytes("rex_quadrotor", 5, 8)
    # SURVEY 8d worked example: F_newton ~ 3.6e5 with the survey's hand count; ours uses the recounted F_dyn
    assert 2.0e5 < fl["newton"] < 4.0e5 and fl["bwd"] > 0 and fl["dual"] > 0
    assert by["newton_launch"] > 0
    assert bench.F_DYN["rex_quadrotor"] == 980 and bench.F_DYN["pendulum"] == 50
