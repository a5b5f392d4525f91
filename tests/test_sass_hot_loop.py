"""CPU: the member decoder's symbol loop must not touch local memory.  ptxas allocates registers across the
__noinline__ calls around the loop, and during development unrelated edits to those functions repeatedly spilled loop
invariants (LUT addresses, the window limit) INTO the loop: the same source then ran between 15.6 and 24.8 ms per GiB
(DESIGN.md 3.1).  Since round 2 the loop is a PTX block; this test disassembles the built object and looks at it."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "deflate-library-java_b200", "build", "inflate.cu.o")


def _sass(kernel):
    out = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True, check=True).stdout
    lines, on = [], False
    for ln in out.splitlines():
        if "Function :" in ln:
            on = kernel in ln
        elif on:
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*/\*", ln)
            if m:
                lines.append(m.group(2).rstrip(" ;"))
    return lines


@pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(OBJ), reason="needs the built object and cuobjdump")
@pytest.mark.parametrize("kernel", ["14inflate_kernelILb0E", "14inflate_kernelILb1E", "20inflate_units_kernel", "19stream_units_kernel"])
def test_symbol_loop_of_the_member_decoder_has_no_local_memory_traffic(kernel):
    """Every kernel carries its own clone of the symbol loop (member decoder, block-parallel decoder of our own streams,
    units of a foreign stream; the member decoder once more with input streaming), allocated in that kernel's context: all
    are checked.  The loop is a PTX block with uniform branches (decode_block_fast, HOT_LOOP): besides local memory it
    must be free of convergence barriers, and stay as short as it was written."""
    sass = _sass(kernel)
    # the literal path: `sh += e >> 27` is the only LEA.HI with a 5-bit shift; the lookup precedes it
    hits = [i for i, ins in enumerate(sass) if re.match(r"LEA\.HI R\d+, R\d+(\.reuse)?, R\d+, RZ, 0x5$", ins)]
    assert hits, "symbol loop not found"
    start = hits[0] - 3
    # the loop (three copies of it: the window registers change roles on a refill instead of being moved) and its exits end
    # where the pairs that need care load their event code (EV_PAIR, 4)
    sts64 = next(i for i in range(start, len(sass)) if sass[i].startswith("STS.64"))
    end = next(i for i in range(sts64, len(sass)) if re.match(r"(@!?P\d\s+)?(IMAD\.MOV\.U32|MOV) R\d+, (RZ, RZ, )?0x4$", sass[i]))
    body = sass[start:end]
    assert 180 < len(body) < 270, len(body)
    # per copy: the table lookups (lit/len, distance), three refills from the line of input in shared memory, the literal
    # store and the queue store are there ...
    assert sum("LDS R" in ins for ins in body) == 15 and sum("STS.U8" in ins for ins in body) == 3
    assert sum(ins.startswith("STS.64") for ins in body) == 3
    # ... and no local memory, no global load, no shuffle, no convergence barrier
    bad = [ins for ins in body if re.search(r"\b(LDL|STL|LDG|SHFL|BSSY|BSYNC|BREAK)\b", ins)]
    assert not bad, bad


def _resources(obj):
    out = subprocess.run(["cuobjdump", "--dump-resource-usage", obj], capture_output=True, text=True, check=True).stdout
    res, name = {}, None
    for ln in out.splitlines():
        m = re.search(r"Function (\S+):", ln)
        if m:
            name = m.group(1)
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", ln)
        if m and name:
            res[name] = dict(zip(("reg", "stack", "shared", "local"), map(int, m.groups())))
    return res


@pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(OBJ), reason="needs the built objects and cuobjdump")
def test_resident_warps_per_sm_of_the_decoders():
    """the 4096 members of config 2 are ONE wave because 7 CTAs of 4 warps fit an SM -- 72 registers per
    thread (7 x 128 x 72 <= 65536) and 33216 B of shared memory per CTA incl. the 1 KiB the driver reserves
    (7 x 33216 <= 233472).  One more register or 200 more bytes and a quarter of the members wait for a second wave.
    No kernel of the library may use local memory beyond its call stack."""
    res = _resources(OBJ)
    decoders = [k for k in res if re.search(r"(14inflate_kernel|20inflate_units_kernel|19stream_units_kernel)", k)]
    assert len(decoders) == 4, sorted(res)
    for k in decoders:
        r = res[k]
        assert r["reg"] * 128 * 7 <= 65536 and r["shared"] * 7 <= 233472, (k, r)
    for name in ("deflate.cu.o", "crc32.cu.o"):
        res.update(_resources(os.path.join(os.path.dirname(OBJ), name)))
    assert len(res) >= 25 and all(r["local"] == 0 for r in res.values()), {k: r for k, r in res.items() if r["local"]}
    # the search kernel owns its SM (one CTA of 1024 threads): 64 registers is the hard limit there
    match = next(r for k, r in res.items() if "12match_kernel" in k)
    assert match["reg"] <= 64, match
