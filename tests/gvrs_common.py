"""Shared by tests/test_gvrs_file.py (CPU) and tests/test_gpu_gvrs_file.py (GPU): the reference's sample files and a
rebuild of a parsed file through the writer (gridfour_b200/gvrs.py)."""
import base64
import json
import os
import struct

HERE = os.path.dirname(os.path.abspath(__file__))


def sample_files():
    doc = json.load(open(os.path.join(HERE, "golden", "gvrs_samples.json")))
    return {k: base64.b64decode(v) for k, v in doc["files"].items()}


def tile_record_parts(body, n_elements):
    """[tileIndex] then [len][bytes] per element (RecordManager.java:386-401)."""
    tile_index = struct.unpack_from("<i", body, 0)[0]
    pos, parts = 4, []
    for _ in range(n_elements):
        n = struct.unpack_from("<i", body, pos)[0]
        parts.append(body[pos + 4:pos + 4 + n])
        pos += 4 + n
    return tile_index, parts


def rebuild(img, context, gpu_tiles):
    """Writes the parsed file again: header and directories from the model, metadata records from their parsed fields,
    tile records re-framed (on the GPU when gpu_tiles and the raster has one element).  context=None: no checksums."""
    from gridfour_b200 import gvrs

    w = gvrs.GvrsWriter(img.spec, uuid=img.uuid, time_modified=img.time_modified, version=img.version)
    recs = img.records
    i = 0
    order = []
    while i < len(recs):
        r = recs[i]
        if r.type_code == gvrs.RECORD_METADATA:
            rd = gvrs._Reader(r.body, 0)
            name = rd.utf()
            record_id, type_code, n = rd.take("<iB3xi")
            content = rd.raw(n)
            description = rd.utf()
            w.add_metadata(name, record_id, type_code, content, description, size=r.size)
            i += 1
        elif r.type_code == gvrs.RECORD_TILE:
            j = i
            while j < len(recs) and recs[j].type_code == gvrs.RECORD_TILE:
                j += 1
            run = recs[i:j]
            if gpu_tiles and context is not None and len(img.spec.elements) == 1:
                arena, offsets, lens, index = bytearray(), [], [], []
                for t in run:
                    ti, parts = tile_record_parts(t.body, 1)
                    arena += bytes((-len(arena)) & 7)
                    offsets.append(len(arena))
                    lens.append(len(parts[0]))
                    index.append(ti)
                    arena += parts[0]
                w.add_tile_records(context, bytes(arena) + bytes(8), offsets, lens, tile_index=index)
            else:
                for t in run:
                    ti, parts = tile_record_parts(t.body, len(img.spec.elements))
                    w.add_tile_record_host(ti, parts, size=t.size)
            i = j
        elif r.type_code == gvrs.RECORD_METADATA_DIR:
            order.append("metadata")
            i += 1
        elif r.type_code == gvrs.RECORD_TILE_DIR:
            order.append("tile")
            i += 1
        else:
            raise AssertionError("unexpected record type %d in a sample file" % r.type_code)
    extended = img.pos_tile_dir != 0 and img.image[img.pos_tile_dir + 1] != 0
    return w.finish(context, extended_directory=extended, directory_order=tuple(order))


def mask_checksums(image, img):
    """The image with every record's checksum field zeroed."""
    b = bytearray(image)
    for pos, size in [(16, img.header_size)] + [(r.pos, r.size) for r in img.records]:
        b[pos + size - 4:pos + size] = bytes(4)
    return bytes(b)
