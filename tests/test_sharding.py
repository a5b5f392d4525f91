"""CPU: the N > 1 host path (unit partitioning, size all-gather, payload gather to rank 0, CRC combine) at
world_size 2 on the gloo backend."""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_unit_range_partitions_exactly():
    import b2d_loader
    b2d_loader.load()
    from importlib import import_module
    sharding = import_module("b2deflate.sharding")
    for n in (0, 1, 7, 8, 1024, 8191):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = sharding.unit_range(n, r, world)
                assert lo == prev and hi >= lo and hi - lo in (n // world, n // world + 1)
                prev = hi
            assert prev == n


def test_gather_to_rank0_world2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = tmp_path / "res.json"
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "dist_worker.py"), str(out)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads(out.read_text())
    assert res["ok"] and res["crc_ok"]
    assert res["n_sizes"] == 37 and res["sum_sizes"] == res["stream_len"]
    assert res["ranges"] == [[0, 19], [19, 37]]
