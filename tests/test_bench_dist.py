"""CPU tests of bench.py's multi-rank plumbing (world_size 2, gloo): barrier, max-over-ranks timing, whole-job
aggregation, and the rule that only rank 0 prints.  The data path has no collective -- frames are independent --
so this is all the N>1 logic there is."""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def torchrun(nproc, *bench_args, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), os.path.join(ROOT, "bench.py")] + list(bench_args)
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def json_lines(out):
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def test_two_ranks_gloo_aggregate_max_over_ranks():
    r = torchrun(2, "--gpus", "2", "--steps", "4", "--warmup", "3", "--backend", "gloo", "--dry-run", "--batch", "10")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = json_lines(r.stdout)
    assert len(lines) == 1                      # rank 0 alone prints
    d = lines[0]
    assert d["n_gpus"] == 2 and d["units_all_ranks"] == 20
    assert d["ms_max"] == 2 * 4                 # slowest rank: (rank+1) ms per step
    assert abs(d["value"] - 20 * 4 / 0.008) < 1e-6


def test_single_process_dry_run():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--dry-run", "--batch", "7", "--steps", "3"],
                       capture_output=True, text=True, cwd=ROOT, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json_lines(r.stdout)[0]
    assert d["n_gpus"] == 1 and d["units_all_ranks"] == 7


def test_reference_arm_under_torchrun_prints_once():
    r = torchrun(2, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--workload", "yuv420p_rgb24_pal")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = json_lines(r.stdout)
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["cores"] >= 1
