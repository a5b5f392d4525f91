"""GPU, >= 2 devices: several GPUs behind the C ABI (b2d_init_devices).  The host entry points shard their units over
the devices, one host thread per GPU, and must return what one GPU returns, byte for byte (host-side
partitioning, sizes scanned, payloads joined at the scanned offsets)."""
import os
import random
import subprocess
import sys
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys, zlib, random
sys.path.insert(0, sys.argv[1])
import numpy as np
import b2d_loader
b2d = b2d_loader.load()
n_dev = int(sys.argv[2])
res = {}
data = np.concatenate([b2d.corpus("mixed", 0xDEF1A7E + k, 16 << 20) for k in range(20)])[:(300 << 20) + 12345]
rng = random.Random(5)
members = []
for i in range(700):
    d = b2d.corpus("text", 1000 + i, rng.randrange(1, 200000)).tobytes()
    c = zlib.compressobj(rng.choice([1, 6, 9]), zlib.DEFLATED, -15)
    members.append(c.compress(d) + c.flush())
members[13] = members[13][:50]              # a truncated member
members[400] = b"\x07garbage"               # reserved block type
out = {}
for tag, devs in (("one", [0]), ("all", list(range(n_dev)))):
    assert b2d.init_devices(devs) == len(devs)
    comp, crc, idx, bits = b2d.deflate_chunks_indexed(data, b2d.make_opts(), crc=0)
    comp2, crc2, idx2 = b2d.deflate_chunks(data, b2d.make_opts(), crc=0)
    dec, dcrc, dst = b2d.inflate_chunks(comp, idx, bits, data.size)
    outs, out_len, cons, crcs, status = b2d.inflate_batch(members, 200000, b2d.INFLATE_CRC32)
    buf = b2d.PinnedBuffer(200000 * len(members))
    outs_p, out_len_p, cons_p, crcs_p, status_p = b2d.inflate_batch(members, 200000, b2d.INFLATE_CRC32, out=buf.array, pinned_in=True)
    out[tag] = dict(comp=bytes(comp), crc=crc, idx=idx.tolist(), bits=bits.tolist(), comp2=bytes(comp2), crc2=crc2,
                    dec_ok=bool(np.array_equal(dec, data)), dst=dst.tolist(), dcrc=dcrc.tolist(), outs=outs,
                    out_len=out_len.tolist(), cons=cons.tolist(), crcs=crcs.tolist(), status=status.tolist(),
                    pinned_same=bool(outs_p == outs and np.array_equal(status_p, status) and np.array_equal(crcs_p, crcs)))
    b2d.shutdown()
a, b = out["one"], out["all"]
res["deflate_same"] = a["comp"] == b["comp"] and a["idx"] == b["idx"] and a["bits"] == b["bits"] and a["crc"] == b["crc"]
res["deflate_plain_same"] = a["comp2"] == b["comp2"] and a["crc2"] == b["crc2"] and a["comp2"] == a["comp"]
res["crc_ok"] = a["crc"] == zlib.crc32(data.data)
res["zlib_ok"] = zlib.decompress(b["comp"], -15) == data.tobytes()
res["inflate_chunks_ok"] = a["dec_ok"] and b["dec_ok"] and not any(b["dst"]) and a["dcrc"] == b["dcrc"]
res["inflate_batch_same"] = all(a[k] == b[k] for k in ("outs", "out_len", "cons", "crcs", "status"))
res["pinned_same"] = a["pinned_same"] and b["pinned_same"]
res["bad_members"] = [b["status"][13], b["status"][400]]
os.environ["B2D_MULTI_GATHER"] = "peer"                 # payloads through GPU 0 (cudaMemcpyPeerAsync), same bytes
assert b2d.init_devices(list(range(n_dev))) == n_dev
comp3, crc3, idx3 = b2d.deflate_chunks(data, b2d.make_opts(), crc=0)
res["peer_gather_same"] = bytes(comp3) == a["comp"] and crc3 == a["crc"]
res["launches"] = b2d.kernel_launches()
b2d.shutdown()
print(json.dumps(res))
'''


def _device_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.skipif(_device_count() < 2, reason="needs at least two GPUs")
def test_all_devices_give_the_bytes_of_one_device(tmp_path):
    """Runs in a process of its own (the session's library is bound to cuda:0 by the b2d fixture)."""
    import json
    n = min(_device_count(), 8)
    r = subprocess.run([sys.executable, "-c", WORKER, ROOT, str(n)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["deflate_same"] and res["deflate_plain_same"] and res["crc_ok"] and res["zlib_ok"], res
    assert res["inflate_chunks_ok"] and res["inflate_batch_same"] and res["pinned_same"] and res["peer_gather_same"], res
    assert res["bad_members"] == [1, 2], res            # UNEXPECTED_END_OF_STREAM, RESERVED_BLOCK_TYPE
    assert res["launches"] > 0
