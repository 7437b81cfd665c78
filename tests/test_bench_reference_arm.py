"""bench.py --impl reference (the CPU arm): contract checks that need no GPU.

Under torchrun every rank is started; rank 0 alone times the unmodified reference (oracle/_ref; the oracle port where that
copy is absent) with ALL host threads -- torchrun exports OMP_NUM_THREADS=1 -- and prints the one JSON line; the other
ranks exit 0 without work.  The line names the same workload as the GPU arm's."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_under_torchrun_prints_one_line_from_rank0_with_all_host_threads():
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["gpu_launches"] == 0
    assert d["metric"] == "train samples/sec" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"]["workload"] == bench.WORKLOAD                      # the GPU arm's line names the same workload
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] == d["value"] and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    assert cb["cores"] == cores, (cb["cores"], cores)                     # not torchrun's OMP_NUM_THREADS=1
    if os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "models")):
        assert cb["kind"] == "reference"
