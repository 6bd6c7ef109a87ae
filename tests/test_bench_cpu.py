"""bench.py contract on the CPU: the reference arm prints one JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3",
                          "--warmup", "3", "--workload", "cfg1", "--ref-seconds", "3"], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert out.returncode == 0, out.stderr
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env-steps/sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 3
    # the reference itself (from /root/reference or its vendored copy oracle/_ref), not the C port
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert "ShmemVecEnv" in d["cpu_baseline"]["sample"] and d["subproc_step_only"]["value"] > d["subproc_step_and_mask"]["value"]
    assert d["port_openmp_full_config"]["kind"] == "port" and d["port_openmp_full_config"]["value"] > d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config("cfg1", 1)            # the same dict our arm prints


def test_reference_arm_other_ranks_exit_quietly():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=60, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_algorithmic_bytes_match_survey():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.algorithmic_bytes((10, 10, 10), False) == 2825
    assert bench.algorithmic_bytes((10, 10, 10), True) == 3225
    assert bench.algorithmic_bytes((20, 20, 20), False) == 11225


def test_ab_picker_on_the_recorded_sessions():
    """tools/ab_pick.py names the build with the best geometric-mean step time over the workloads (profiles/r1k_ab_*.jsonl
    are the two A/B sessions of round 1)."""
    import subprocess
    import sys
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    pick = lambda f: subprocess.check_output([sys.executable, os.path.join(root, "tools", "ab_pick.py"),
                                              os.path.join(root, "profiles", f)]).decode().strip()
    assert pick("r1k_ab_first.jsonl") == "packed"        # before the square copy had a single call site
    assert pick("r1k_ab_final.jsonl") == "fullps"        # = today's default build
    assert pick("does_not_exist.jsonl") == "full"


def test_vendored_reference_is_unmodified():
    """oracle/make_ref.py copies the reference's envs/ acktr/ baselines/ packages byte for byte into oracle/_ref (git-ignored);
    the manifest's SHA-256 of every file must match the copy and, where the reference tree is present, the source."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_ref
    dest = make_ref.build()
    if dest is None:
        import pytest
        pytest.skip("neither /root/reference nor an earlier oracle/_ref")
    assert make_ref.verify()
    import json
    files = json.load(open(os.path.join(dest, "MANIFEST.json")))["files"]
    assert "acktr/envs.py" in files and "envs/bpp0/bin3D.py" in files and "baselines/common/vec_env/shmem_vec_env.py" in files
    # nothing of it is tracked by git: reference sources stay out of this repository's history
    tracked = subprocess.run(["git", "ls-files", "oracle/_ref"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    assert tracked == ""
