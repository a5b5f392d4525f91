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


def test_slice_ranges_are_slice_major_and_cover_every_unit_once():
    """bench.py config4_leg relies on it: concatenating, slice after slice, the ranks' sub-ranges in rank order gives
    0 .. n-1 in stream order, so a slice's payloads can go to their final offsets once its sizes are known."""
    import b2d_loader
    b2d_loader.load()
    from importlib import import_module
    sharding = import_module("b2deflate.sharding")
    for n in (0, 1, 5, 41, 1024, 8192):
        for world in (1, 2, 3, 4, 8):
            for n_slices in (1, 2, 3, 4):
                per_rank = [sharding.slice_ranges(n, r, world, n_slices) for r in range(world)]
                assert all(len(p) == n_slices for p in per_rank)
                order = []
                for k in range(n_slices):
                    for r in range(world):
                        lo, hi = per_rank[r][k]
                        assert hi >= lo
                        order += list(range(lo, hi))
                assert order == list(range(n))
    # config 4 itself: 8192 chunks, 8 ranks, 2 slices of 512 chunks per rank
    assert [hi - lo for lo, hi in sharding.slice_ranges(8192, 7, 8, 2)] == [512, 512]


def _torchrun(tmp_path, nproc, *extra):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = tmp_path / "res.json"
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "dist_worker.py"), str(out), *extra],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads(out.read_text())


def test_pipelined_slice_gather_world2_and_world3(tmp_path):
    """The exchange of config 4 (slice-major ranges, per-slice size all-gather, payloads to their final offsets on
    rank 0) over gloo: the gathered stream is one DEFLATE stream, and the gathered index finds every chunk in it."""
    for nproc, slices in ((2, 3), (3, 2)):
        res = _torchrun(tmp_path, nproc, str(slices))
        assert res["world"] == nproc and res["owned_once"]
        assert res["ok"] and res["per_chunk_ok"] and res["crc_ok"]
        assert res["n_sizes"] == 41 and res["sum_sizes"] == res["stream_len"]


def test_gather_to_rank0_world2(tmp_path):
    res = _torchrun(tmp_path, 2)
    assert res["ok"] and res["crc_ok"]
    assert res["n_sizes"] == 37 and res["sum_sizes"] == res["stream_len"]
    assert res["ranges"] == [[0, 19], [19, 37]]
