"""N>1 host logic on CPU: two gloo ranks shard one Fock build and combine it with a single all-reduce."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port())] + args
    env = dict(os.environ, OMP_NUM_THREADS="2")
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=env)


def test_sharded_build_allreduce_world2(orc):
    r = _torchrun([os.path.join(ROOT, "tests", "_gloo_worker.py")])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "GLOO_OK" in r.stdout


def test_reference_arm_under_torchrun_prints_one_line(orc):
    import json

    r = _torchrun([os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                   "--cpu-seconds", "2", "--workload", "h2o4_631gs"])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "fock_builds_per_s" and j["value"] > 0
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["cpu_baseline"]["kind"] == "port"
