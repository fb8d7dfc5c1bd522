"""Debug aid: one tile through CodecHuffman.decode on the GPU vs the oracle; prints where the first difference is."""
import sys
import numpy as np
sys.path.insert(0, ".")
import gridfour_b200 as g4
from oracle import g4oracle as o

for (r, c) in ((90, 120), (180, 240), (45, 60)):
    t = o.terrain_i32(0, 0, r, c)
    p, pred = o.codec_encode_i32(o.CODEC_HUFFMAN, 0, t)
    out = g4.CodecHuffman().decode(r, c, p)
    d = np.argwhere(out != t)
    print(r, c, "pred", pred, "len", len(p), "nM32", int.from_bytes(p[6:10], "little"), "ndiff", len(d), d[:5].tolist())
    if len(d):
        rr, cc = d[0]
        print(" got", out[rr, max(0, cc - 2):cc + 6], " want", t[rr, max(0, cc - 2):cc + 6])
