"""Debug helper (round 2): decode a config-3 band, report mismatching tiles / cells and per-kernel times."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gridfour_b200 as g4
from oracle import g4oracle as oracle

tr, tc = 180, 240
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 360
grid = oracle.terrain_i32(int(os.environ.get("ROW0", "0")), 0, tr * rows, tc * cols, n_threads=16)
spec = g4.CodecSpecification(default=False)
if os.environ.get("THREE"):
    spec.addCompressionCodec("GvrsHuffman", g4.CodecHuffman, g4.CodecHuffman)
    spec.addCompressionCodec("GvrsDeflate", g4.CodecDeflate, g4.CodecDeflate)
spec.addCompressionCodec("LSOP12", g4.LsEncoder12, g4.LsDecoder12)
master = g4.CodecMaster(spec)
batch = master.encodeTiles(grid, tr, tc)
out = master.decodeTiles(batch)
bad = out != grid
print("mismatching cells:", int(bad.sum()), "of", grid.size)
if bad.any():
    tiles = {}
    rr, cc = np.nonzero(bad)
    for r, c in zip(rr[:200000], cc[:200000]):
        tiles.setdefault((r // tr, c // tc), []).append((r % tr, c % tc))
    print("tiles with mismatches:", len(tiles))
    for k, v in list(tiles.items())[:6]:
        v = sorted(v)
        print(" tile", k, "n=", len(v), "first cells", v[:6], "rows", sorted(set(r for r, _ in v))[:8], "cols", sorted(set(c for _, c in v))[:12])
        r, c = v[0]
        R0, C0 = k[0] * tr, k[1] * tc
        print("   got ", out[R0 + r, C0 + max(0, c - 2):C0 + c + 6], "\n   want", grid[R0 + r, C0 + max(0, c - 2):C0 + c + 6])
