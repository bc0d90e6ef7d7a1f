"""bench.py contract on the CPU: the reference arm (the only arm that runs without a GPU) prints exactly one JSON
line with the keys the driver reads; our arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, env=e, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--workload", "C2", "--cpu-scale", "20", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "lsqr_effective_hbm_GBps" and d["unit"] == "GB/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["iters_per_s"] > 0 and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and "C2 at 1/20 scale" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "C2"
    assert cb["cpu_model"] and cb["cores_present"] >= 1 and cb["sample_scale"] == 20
    assert "-ffp-contract=off" in cb["sample"]


def test_reference_arm_never_maps_the_engine_library():
    """The reference arm times the oracle port only: the process must not load liblsqr_b200.so (the driver records
    which of the repo's .so files each arm maps)."""
    code = ("import sys, os; sys.path.insert(0, %r); sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'C2', "
            "'--cpu-scale', '50', '--steps', '1', '--warmup', '0']; import bench; bench.main()\n"
            "maps = open('/proc/self/maps').read()\n"
            "sys.stderr.write('ENGINE_MAPPED=%%d ORACLE_MAPPED=%%d' %% ('liblsqr_b200' in maps, 'liblsqr_oracle' in maps))\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "ENGINE_MAPPED=0 ORACLE_MAPPED=1" in r.stderr
    assert "lsqr_b200" not in [m.split(".")[0] for m in r.stderr.split() if m.startswith("IMPORTED:")]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_default_workload_is_the_north_star_problem_on_every_n():
    # in a subprocess: importing bench.py points file descriptor 1 at stderr for the life of the process
    code = ("import sys; sys.path.insert(0, %r); sys.argv = ['bench.py']; import bench\n"
            "out = []\n"
            "for a in (['--gpus', '1'], ['--gpus', '2'], ['--gpus', '4'], ['--gpus', '8'], ['--workload', 'C3']):\n"
            "    sys.argv = ['bench.py'] + a; out.append(bench.pick_workload(bench.parse_args()))\n"
            "sys.stderr.write('PICK ' + ' '.join(out))\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "PICK C5 C5 C5 C5 C3" in r.stderr
