#!/usr/bin/env python3
"""Extracts the tile payloads of the reference's own binary fixtures into tests/golden/fixtures.json.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py

Source: core/src/test/resources/org/gridfour/gvrs/SampleFiles/*.gvrs, contents documented in
SampleFiles/README.txt:8-55.  File layout walked here: 16-byte preamble, header record at offset 16
([size:int32][type:byte=6]...), then 8-byte aligned records [size:int32][type:byte][3 reserved]
(C/gvrs/RecordManager.java:70-78); a tile record (type 2) holds [tileIndex:int32] then one
[len:int32][bytes] block per element (RecordManager.java:386-401, RasterTile.java:243-253).
Only the element payload bytes are stored (hex) -- they are data produced by the reference's Java
codecs (JDK zlib, HuffmanEncoder, LsEncoder) and are the decode-side golden vectors.
"""
import json, os, struct

SRC = "/root/reference/core/src/test/resources/org/gridfour/gvrs/SampleFiles"
HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (grid rows, grid cols, tile rows, tile cols, element kind, codec id list in file order)
SAMPLES = {
    "Sample01_IntNoComp": (10, 10, 5, 5, "int", []),
    "Sample04_ShortComp": (100, 100, 50, 50, "short", ["GvrsHuffman", "GvrsDeflate", "GvrsFloat"]),
    "Sample05_IntComp": (100, 100, 50, 50, "int", ["GvrsHuffman", "GvrsDeflate", "GvrsFloat"]),
    "Sample06_FltComp": (100, 100, 50, 50, "float", ["GvrsHuffman", "GvrsDeflate", "GvrsFloat"]),
    "Sample07_ICFComp": (100, 100, 50, 50, "icf", ["GvrsHuffman", "GvrsDeflate", "GvrsFloat"]),
    "Sample14_LSOP": (101, 101, 101, 101, "icf1000", ["LSOP12"]),
}


def tiles_of(path):
    b = open(path, "rb").read()
    assert b[:11] == b"gvrs raster"
    pos = 16 + struct.unpack_from("<i", b, 16)[0]
    out = {}
    while pos + 8 <= len(b):
        size, typ = struct.unpack_from("<iB", b, pos)
        if size <= 0:
            break
        if typ == 2:
            tile_index, ln = struct.unpack_from("<ii", b, pos + 8)
            out[tile_index] = b[pos + 16 : pos + 16 + ln]
        pos += size
    return out


def main():
    doc = {"source": "gridfour core/src/test/resources/org/gridfour/gvrs/SampleFiles", "samples": {}}
    for name, (gr, gc, tr, tc, kind, codecs) in SAMPLES.items():
        tiles = tiles_of(os.path.join(SRC, name + ".gvrs"))
        doc["samples"][name] = {
            "grid_rows": gr, "grid_cols": gc, "tile_rows": tr, "tile_cols": tc, "kind": kind, "codecs": codecs,
            "tiles": {str(k): v.hex() for k, v in sorted(tiles.items())},
        }
    with open(os.path.join(HERE, "fixtures.json"), "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote fixtures.json:", {k: len(v["tiles"]) for k, v in doc["samples"].items()})


if __name__ == "__main__":
    main()
