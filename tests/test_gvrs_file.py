"""not-gpu: the file layer around the codec path against the reference's own sample files
(tests/golden/gvrs_samples.json, written by the reference's Java code): the oracle's CRC-32C against every stored record
checksum, and a parse -> rebuild of every file that must reproduce its bytes (checksum fields aside: the product computes
them on the GPU, tests/test_gpu_gvrs_file.py)."""
import struct

import numpy as np
import pytest

from gvrs_common import mask_checksums, rebuild, sample_files, tile_record_parts


@pytest.fixture(scope="module")
def oracle():
    from oracle import g4oracle

    g4oracle.build()
    return g4oracle


def test_oracle_crc32c_matches_every_record_checksum_of_the_reference_files(oracle):
    from gridfour_b200 import gvrs

    checked = 0
    for name, image in sample_files().items():
        img = gvrs.GvrsImage.parse(image)
        if not img.spec.checksum:
            continue
        off, size, stored = img.record_ranges()
        for o, s, c in zip(off, size, stored):
            assert oracle.crc32c(image[int(o):int(o) + int(s)]) == int(c), (name, int(o))
            checked += 1
    assert checked >= 120
    assert oracle.crc32c(b"123456789") == 0xE3069283  # RFC 3720 check value


def test_parse_and_rebuild_reproduces_the_reference_files():
    from gridfour_b200 import gvrs

    files = sample_files()
    assert len(files) == 17
    for name, image in files.items():
        img = gvrs.GvrsImage.parse(image)
        again = rebuild(img, None, gpu_tiles=False)
        assert len(again) == len(image), name
        assert again == mask_checksums(image, img), name


def test_specification_and_directory_of_known_samples():
    from gridfour_b200 import gvrs

    files = sample_files()
    img = gvrs.GvrsImage.parse(files["Sample05_IntComp.gvrs"])
    s = img.spec
    assert (s.n_rows, s.n_cols, s.tile_rows, s.tile_cols) == (100, 100, 50, 50)
    assert s.codecs == ["GvrsHuffman", "GvrsDeflate", "GvrsFloat"] and s.checksum
    assert [e.type_code for e in s.elements] == [gvrs.ELEM_INTEGER] and s.elements[0].fill_value == -2147483648
    d = img.tile_directory()
    assert sorted(d) == [0, 1, 2, 3]
    for t, pos in d.items():  # the directory points at the content of the record that carries the tile index
        assert struct.unpack_from("<i", image_of(files, "Sample05_IntComp.gvrs"), pos)[0] == t
    img = gvrs.GvrsImage.parse(files["Sample14_LSOP.gvrs"])
    assert img.spec.codecs == ["LSOP12"] and img.spec.elements[0].type_code == gvrs.ELEM_INT_CODED_FLOAT
    img = gvrs.GvrsImage.parse(files["Sample08_MixedTypes.gvrs"])
    assert len(img.spec.elements) == 2 and not img.spec.checksum
    for r in img.records:
        if r.type_code == gvrs.RECORD_TILE:
            _, parts = tile_record_parts(r.body, 2)
            n = img.spec.tile_rows * img.spec.tile_cols
            assert [len(p) for p in parts] == [e.standard_size(n) for e in img.spec.elements]


def image_of(files, name):
    return files[name]


def test_new_file_layout_without_gpu():
    """Header arithmetic of a file written from scratch: sizes, alignment, directory round trip."""
    from gridfour_b200 import gvrs

    spec = gvrs.GvrsSpec(360, 480, 90, 120, [gvrs.ElementSpec.integer("z")], ["GvrsHuffman", "LSOP12"], checksum=True,
                         product_label="test")
    w = gvrs.GvrsWriter(spec, uuid=bytes(range(16)), time_modified=1234567)
    for name, rid, typ, content, desc in gvrs.codec_metadata(spec.codecs):
        w.add_metadata(name, rid, typ, content, desc)
    payloads = {t: bytes([t]) * (10 + 3 * t) for t in range(16) if t != 5}
    for t, p in payloads.items():
        w.add_tile_record_host(t, [p])
    image = w.finish(None)
    assert len(image) % 8 == 0
    img = gvrs.GvrsImage.parse(image)
    assert img.spec.codecs == spec.codecs and img.spec.product_label == "test" and img.uuid == bytes(range(16))
    d = img.tile_directory()
    assert sorted(d) == sorted(payloads)
    for t, pos in d.items():
        ti, n = struct.unpack_from("<ii", image, pos)
        assert ti == t and image[pos + 8:pos + 8 + n] == payloads[t] and pos % 8 == 0
    assert [r.type_code for r in img.records][-2:] == [gvrs.RECORD_METADATA_DIR, gvrs.RECORD_TILE_DIR]
    ext = gvrs.tile_directory_content(spec, d, extended=True)
    assert ext[1] == 1 and len(ext) == 8 + 16 + 8 * 16
    assert np.all(np.frombuffer(image[16 + img.header_size - 4:16 + img.header_size], dtype=np.uint8) == 0)
