#!/usr/bin/env python3
"""Stores the reference's own GVRS sample files (17 files, 25 KB together) as base64 in tests/golden/gvrs_samples.json.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_gvrs_samples.py

Source: core/src/test/resources/org/gridfour/gvrs/SampleFiles/*.gvrs (written by the reference's Java code; contents
documented in SampleFiles/README.txt).  They are the golden vectors of the file layer: every record's CRC-32C, the record
framing, the header / specification block, the metadata and tile directories (tests/test_gvrs_file.py and
tests/test_gpu_gvrs_file.py rebuild each file byte for byte and decode the tiles straight from the image).
"""
import base64, glob, json, os

SRC = "/root/reference/core/src/test/resources/org/gridfour/gvrs/SampleFiles"
HERE = os.path.dirname(os.path.abspath(__file__))

doc = {"source": "gridfour core/src/test/resources/org/gridfour/gvrs/SampleFiles", "files": {}}
for path in sorted(glob.glob(os.path.join(SRC, "*.gvrs"))):
    doc["files"][os.path.basename(path)] = base64.b64encode(open(path, "rb").read()).decode("ascii")
json.dump(doc, open(os.path.join(HERE, "gvrs_samples.json"), "w"), indent=1)
print("wrote gvrs_samples.json:", len(doc["files"]), "files")
