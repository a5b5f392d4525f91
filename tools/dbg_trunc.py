import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import random, zlib
from oracle import oracle as O
from util import *
import b2d_loader
b2d=b2d_loader.load()
import os
if os.environ.get("DBGSO"): b2d.binding.SO_PATH=os.environ["DBGSO"]
b2d.init(0)
def _text(rng, n):
    words = [bytes(rng.choices(b"etaoinshrdlucmfw", k=rng.randrange(1, 10))) for _ in range(700)]
    out = bytearray()
    while len(out) < n:
        out += rng.choice(words) + b" "
    return bytes(out[:n])
rng = random.Random(404)
s = zlib_raw(_text(rng, 3000), 6)
full=zlib.decompress(s,-15)
bad=0
for cut in ([int(os.environ['CUT'])] if os.environ.get('CUT') else range(0,len(s))):
    m=s[:cut]
    outs, out_len, consumed, crc, status = b2d.inflate_batch([m], 8192)
    st,out,cons=O.inflate(m,out_cap=8192)
    if outs[0]!=out or status[0]!=st:
        bad+=1
        if bad<6:
            g=outs[0]; fd=next((i for i in range(min(len(g),len(full))) if g[i]!=full[i]),-1)
            print('first diff vs truth',fd,'consumed',consumed[0],cons)
            print(cut, status[0], st, len(outs[0]), len(out), outs[0][-6:], out[-6:], full[len(out)-3:len(out)+6])
print('bad',bad)
