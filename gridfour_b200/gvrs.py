"""GVRS file images either side of the codec path.

Host-side mirror of the reference's file layout -- pure offset arithmetic, no sample data is touched on the CPU:
  * C/gvrs/GvrsFile.java:241-300,560-623        preamble, header record, fixed header slots, close() sequence
  * C/gvrs/GvrsFileSpecification.java:1170-1285  specification block (write) and :960-1045 (read)
  * C/gvrs/RecordManager.java:70-78,137-139,161-204,386-490,835-883,991-1017   records, tile records, directories
  * C/gvrs/TileDirectory.java:244-278            tile directory (int32 = content position / 8)
  * C/gvrs/GvrsMetadata.java:217-234             metadata record content
  * C/io/BufferedRandomAccessFile.java:689-701   leWriteUTF = [length:uint16 LE][UTF-8 bytes]
(C/ = /root/reference/core/src/main/java/org/gridfour/).

Everything that walks sample bytes runs on the GPU through the C ABI: record checksums (g4_crc32c), tile-record framing
(g4_pack_tile_records), record validation + payload location (g4_unpack_tile_records) and the codecs themselves
(g4_encode_tiles / g4_decode_tiles with the file image as the arena).  The batched tile entry points serve rasters with
one element per tile (the BASELINE configurations); files with several elements are parsed and re-framed record by
record with the GPU computing the checksums, and read element by element (read_raster walks the [len][bytes] chains
on the host and hands the payload offsets to g4_decode_tiles).
"""
import ctypes as C
import struct

import numpy as np

from . import _lib
from ._lib import G4_MEM_HOST, check

RECORD_FREESPACE, RECORD_METADATA, RECORD_TILE, RECORD_FREESPACE_DIR, RECORD_METADATA_DIR, RECORD_TILE_DIR, RECORD_HEADER = range(7)
ELEM_INTEGER, ELEM_INT_CODED_FLOAT, ELEM_FLOAT, ELEM_SHORT = 0, 1, 2, 3  # GvrsElementType.java:50-64
_ELEM_BYTES = {ELEM_INTEGER: 4, ELEM_INT_CODED_FLOAT: 4, ELEM_FLOAT: 4, ELEM_SHORT: 2}
FILEPOS_HEADER_RECORD = 16
FILEPOS_FREESPACE_DIR, FILEPOS_METADATA_DIR, FILEPOS_TILE_DIR = 56, 64, 80
METADATA_TYPE_STRING = 9  # GvrsMetadataType.STRING (the type byte of the two codec metadata records in every sample file)


def multiple_of_8(n):
    return (n + 7) & ~7


def _utf(s):
    b = (s or "").encode("utf-8")
    return struct.pack("<H", len(b)) + b


class _Reader:
    def __init__(self, b, pos):
        self.b, self.pos = b, pos

    def take(self, fmt):
        v = struct.unpack_from(fmt, self.b, self.pos)
        self.pos += struct.calcsize(fmt)
        return v if len(v) > 1 else v[0]

    def raw(self, n):
        v = bytes(self.b[self.pos:self.pos + n])
        self.pos += n
        return v

    def utf(self):
        n = self.take("<H")
        return self.raw(n).decode("utf-8")

    def skip_to_4(self):
        self.pos += (-self.pos) & 3


class ElementSpec:
    """GvrsElementSpecification*: type code, name, the type's range block (kept as raw little-endian bytes so that a
    re-serialised header is bit-identical), label, description, unit of measure."""

    def __init__(self, type_code, name, continuous=False, range_block=b"", label="", description="", unit=""):
        self.type_code, self.name, self.continuous = type_code, name, continuous
        self.range_block, self.label, self.description, self.unit = range_block, label, description, unit

    @property
    def fill_value(self):
        if self.type_code == ELEM_SHORT:
            return struct.unpack_from("<h", self.range_block, 4)[0]
        if self.type_code == ELEM_INTEGER:
            return struct.unpack_from("<i", self.range_block, 8)[0]
        if self.type_code == ELEM_FLOAT:
            return struct.unpack_from("<f", self.range_block, 8)[0]
        return struct.unpack_from("<i", self.range_block, 28)[0]  # INT_CODED_FLOAT: the integer fill value

    @staticmethod
    def integer(name, min_value=-(2 ** 31) + 1, max_value=2 ** 31 - 1, fill_value=-(2 ** 31), **kw):
        return ElementSpec(ELEM_INTEGER, name, range_block=struct.pack("<iii", min_value, max_value, fill_value), **kw)

    @staticmethod
    def floating(name, min_value=-3.4028235e38, max_value=3.4028235e38, fill_value=float("nan"), **kw):
        return ElementSpec(ELEM_FLOAT, name, range_block=struct.pack("<fff", min_value, max_value, fill_value), **kw)

    def standard_size(self, n_cells):
        n = n_cells * _ELEM_BYTES[self.type_code]
        return (n + 3) & ~3  # TileElement.java:89-93: the 2-byte form is padded to a multiple of 4


_RANGE_BYTES = {ELEM_SHORT: 6, ELEM_FLOAT: 12, ELEM_INT_CODED_FLOAT: 32, ELEM_INTEGER: 12}


class GvrsSpec:
    """The part of GvrsFileSpecification that is stored in the header.  The georeferencing block (raster space code,
    coordinate system code, 18 doubles) is opaque to the codec path and kept as raw bytes."""

    def __init__(self, n_rows, n_cols, tile_rows, tile_cols, elements, codecs=(), checksum=False, product_label="",
                 georef=None):
        self.n_rows, self.n_cols, self.tile_rows, self.tile_cols = n_rows, n_cols, tile_rows, tile_cols
        self.elements, self.codecs, self.checksum, self.product_label = list(elements), list(codecs), bool(checksum), product_label
        if georef is None:  # what GvrsFileSpecification's constructor sets up: unit cells, identity transforms
            d = [0.0, 0.0, n_cols - 1.0, n_rows - 1.0, 1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0]
            georef = bytes([0, 0, 0, 0, 0, 0, 0]) + struct.pack("<18d", *d)
        self.georef = georef  # rasterSpace, coordinateSystem, 5 reserved bytes, 18 doubles = 151 bytes

    @property
    def tiles_down(self):
        return (self.n_rows + self.tile_rows - 1) // self.tile_rows

    @property
    def tiles_across(self):
        return (self.n_cols + self.tile_cols - 1) // self.tile_cols

    @property
    def standard_tile_bytes(self):
        n = self.tile_rows * self.tile_cols
        return sum(e.standard_size(n) for e in self.elements)

    def serialize(self, file_pos):
        """GvrsFileSpecification.write (:1170-1285); file_pos = absolute position of the first byte (padding to multiples
        of 4 is relative to the file position, padMultipleOf4 :1158-1164)."""
        out = bytearray()
        out += struct.pack("<iiiiii", self.n_rows, self.n_cols, self.tile_rows, self.tile_cols, 0, 0)
        out += bytes([1 if self.checksum else 0]) + self.georef
        out += struct.pack("<i", len(self.elements))
        for e in self.elements:
            out += bytes([e.type_code, 1 if e.continuous else 0]) + bytes(6) + _utf(e.name)
            out += bytes((-(file_pos + len(out))) & 3)
            out += e.range_block + _utf(e.label) + _utf(e.description) + _utf(e.unit)
            out += bytes((-(file_pos + len(out))) & 3)
        out += struct.pack("<i", len(self.codecs))
        for c in self.codecs:
            out += _utf(c)
        out += _utf(self.product_label)
        return bytes(out)

    @staticmethod
    def parse(b, pos):
        r = _Reader(b, pos)
        n_rows, n_cols, tile_rows, tile_cols, _, _ = r.take("<iiiiii")
        checksum = r.take("<B") != 0
        georef = r.raw(7 + 18 * 8)
        elements = []
        for _ in range(r.take("<i")):
            type_code, continuous = r.take("<BB")
            r.raw(6)
            name = r.utf()
            r.skip_to_4()
            if type_code not in _RANGE_BYTES:
                raise IOError("Unsupported value for data-type code: %d" % type_code)
            block = r.raw(_RANGE_BYTES[type_code])
            label, description, unit = r.utf(), r.utf(), r.utf()
            r.skip_to_4()
            elements.append(ElementSpec(type_code, name, continuous != 0, block, label, description, unit))
        codecs = [r.utf() for _ in range(r.take("<i"))]
        label = r.utf()
        return GvrsSpec(n_rows, n_cols, tile_rows, tile_cols, elements, codecs, checksum, label, georef), r.pos


class Record:
    def __init__(self, pos, size, type_code, body):
        self.pos, self.size, self.type_code, self.body = pos, size, type_code, body  # body = bytes 8 .. size-4 (content + padding)

    @property
    def content_pos(self):
        return self.pos + 8


class GvrsImage:
    """A parsed GVRS file image (version 1.3 and later).  `records` lists every record after the header in file order."""

    def __init__(self):
        self.version = (1, 4)
        self.uuid = bytes(16)
        self.time_modified = 0
        self.time_opened = 0
        self.spec = None
        self.records = []
        self.header_size = 0
        self.pos_freespace_dir = self.pos_metadata_dir = self.pos_tile_dir = 0
        self.image = b""

    # ---- reading ---------------------------------------------------------------------------------------------------
    @staticmethod
    def parse(image):
        b = bytes(image)
        if b[:11] != b"gvrs raster" or len(b) < 112:
            raise IOError("not a GVRS raster file")
        g = GvrsImage()
        g.image = b
        g.version = (b[12], b[13])
        if g.version < (1, 3):
            raise IOError("GVRS version %d.%d is not supported (1.3 and later)" % g.version)
        g.header_size, htype = struct.unpack_from("<iB", b, FILEPOS_HEADER_RECORD)
        if htype != RECORD_HEADER or g.header_size < 96 or FILEPOS_HEADER_RECORD + g.header_size > len(b):
            raise IOError("damaged GVRS header record")
        g.uuid = b[24:40]
        g.time_modified, g.time_opened = struct.unpack_from("<qq", b, 40)
        g.pos_freespace_dir, g.pos_metadata_dir = struct.unpack_from("<qq", b, FILEPOS_FREESPACE_DIR)
        g.pos_tile_dir = struct.unpack_from("<q", b, FILEPOS_TILE_DIR)[0]
        g.spec, _ = GvrsSpec.parse(b, 104)
        pos = FILEPOS_HEADER_RECORD + g.header_size
        while pos + 8 <= len(b):
            size, type_code = struct.unpack_from("<iB", b, pos)
            if size < 16 or (size & 7) or pos + size > len(b) or type_code > RECORD_TILE_DIR:
                raise IOError("damaged record at file position %d" % pos)
            g.records.append(Record(pos, size, type_code, b[pos + 8:pos + size - 4]))
            pos += size
        return g

    def tile_directory(self):
        """{tileIndex: content position}; RecordManager.readTileDirectory (:835-858) + TileDirectory.readTilePositions."""
        if self.pos_tile_dir == 0:
            return {}
        b, p = self.image, self.pos_tile_dir
        extended = b[p + 1] != 0
        row0, col0, n_rows, n_cols = struct.unpack_from("<iiii", b, p + 8)
        out = {}
        q = p + 24
        for i in range(n_rows):
            for j in range(n_cols):
                if extended:  # TileDirectoryExtended: 64-bit positions
                    off = struct.unpack_from("<q", b, q)[0]
                    q += 8
                else:
                    off = struct.unpack_from("<I", b, q)[0] * 8
                    q += 4
                if off:
                    out[(row0 + i) * self.spec.tiles_across + (col0 + j)] = off
        return out

    def record_ranges(self, include_header=True):
        """(offsets, sizes) of the checksummed part of every record: the first size-4 bytes."""
        recs = ([(FILEPOS_HEADER_RECORD, self.header_size)] if include_header else []) + [(r.pos, r.size) for r in self.records]
        return (np.array([p for p, _ in recs], dtype=np.uint64), np.array([s - 4 for _, s in recs], dtype=np.uint32),
                np.array([struct.unpack_from("<I", self.image, p + s - 4)[0] for p, s in recs], dtype=np.uint32))

    def verify_checksums(self, context):
        """GPU CRC-32C of every record against the stored values.  Returns the number of records checked."""
        if not self.spec.checksum:
            return 0
        off, size, stored = self.record_ranges()
        crc = crc32c_ranges(context, self.image, off, size)
        bad = np.nonzero(crc != stored)[0]
        if len(bad):
            raise IOError("Checksum mismatch in record at file position %d" % int(off[bad[0]]))
        return len(off)

    def locate_payloads(self, ctx, element=0, verify=True):
        """(payload offsets, lengths, status) of one element of every tile, as g4_decode_tiles takes them with the file image
        as its arena; status G4_DECLINED marks tiles the file does not hold."""
        from ._lib import G4_DECLINED

        spec = self.spec
        if not 0 <= element < len(spec.elements):
            raise ValueError("no such element")
        n_tiles = spec.tiles_down * spec.tiles_across
        directory = self.tile_directory()
        pos = np.zeros(n_tiles, dtype=np.uint64)
        for t, p in directory.items():
            pos[t] = p
        if len(spec.elements) == 1:
            payload_off, lens, status = unpack_tile_records(ctx, self.image, pos, checksum=verify and spec.checksum)
            bad = np.nonzero(status < 0)[0]
            if len(bad):
                raise IOError("damaged tile record for tile %d" % int(bad[0]))
        else:
            b = self.image
            payload_off = np.zeros(n_tiles, dtype=np.uint64)
            lens = np.zeros(n_tiles, dtype=np.uint32)
            status = np.full(n_tiles, G4_DECLINED, dtype=np.int32)
            rec_off, rec_len, rec_crc = [], [], []
            for t, p in directory.items():
                size, type_code = struct.unpack_from("<iB", b, p - 8)
                if type_code != RECORD_TILE or size < 24 or (size & 7) or p - 8 + size > len(b):
                    raise IOError("damaged tile record for tile %d" % t)
                q = p + 4
                for k in range(len(spec.elements)):
                    ln = struct.unpack_from("<i", b, q)[0]
                    if ln < 0 or q + 4 + ln > p - 8 + size - 4:
                        raise IOError("damaged tile record for tile %d" % t)
                    if k == element:
                        payload_off[t], lens[t], status[t] = q + 4, ln, 0
                    q += 4 + ln
                rec_off.append(p - 8)
                rec_len.append(size - 4)
                rec_crc.append(struct.unpack_from("<I", b, p - 8 + size - 4)[0])
            if verify and spec.checksum and rec_off:
                crc = crc32c_ranges(ctx, b, rec_off, rec_len)
                if np.any(crc != np.array(rec_crc, dtype=np.uint32)):
                    raise IOError("Checksum mismatch in a tile record")
        return payload_off, lens, status

    def read_raster(self, master, element=0, verify=True, crop=False):
        """crop=True returns the n_rows x n_cols raster without the fill-valued margin of the edge tiles.
        Decodes every tile of one element of the raster on the GPU straight from the file image: the image is the arena
        of g4_decode_tiles.  One-element rasters: record validation, checksum check and payload location by
        g4_unpack_tile_records.  Rasters with several elements per tile: the [len][bytes] chain of every tile record is
        walked on the host (structure only) and the record checksums are checked by g4_crc32c.  Tiles that are absent from
        the file come back filled with the element's fill value (RasterTile.setToNullState)."""
        spec = self.spec
        e = spec.elements[element] if 0 <= element < len(spec.elements) else None
        payload_off, lens, status = self.locate_payloads(master._context(), element, verify)
        dtype = {ELEM_INTEGER: np.int32, ELEM_INT_CODED_FLOAT: np.int32, ELEM_FLOAT: np.float32, ELEM_SHORT: np.int16}[e.type_code]
        grid = master.decodeImageTiles(self.image, payload_off, lens, status, spec.tiles_down, spec.tiles_across, spec.tile_rows,
                                       spec.tile_cols, dtype, e.fill_value)
        return grid[:spec.n_rows, :spec.n_cols] if crop else grid


# ---- GPU-backed primitives -------------------------------------------------------------------------------------------
def _u8(buf):
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    return np.ascontiguousarray(a)


def crc32c_ranges(context, data, offsets, sizes):
    """CRC-32C (GridfourCRC32C) of data[offsets[i] : offsets[i]+sizes[i]] for every i, computed on the GPU."""
    a = _u8(data)
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    sz = np.ascontiguousarray(sizes, dtype=np.uint32)
    out = np.zeros(len(off), dtype=np.uint32)
    if len(off) and int((off + sz).max()) > a.size:
        raise ValueError("range outside the buffer")
    check(_lib.lib().g4_crc32c(context._h, G4_MEM_HOST, a.ctypes.data, off.ctypes.data, sz.ctypes.data, len(off), out.ctypes.data),
          "g4_crc32c")
    return out


def crc32c(context, data):
    return int(crc32c_ranges(context, data, [0], [len(data)])[0])


def pack_tile_records(context, arena, offsets, lens, base_pos, checksum, tile_index=None, first_tile_index=0):
    """RecordManager.writeTile framing of a batch of one-element tile payloads.  Returns (record bytes, content positions)."""
    a = _u8(arena)
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    ln = np.ascontiguousarray(lens, dtype=np.uint32)
    n = len(off)
    L = _lib.lib()
    cap = int(L.g4_tile_records_bound(n, int(ln.sum())))
    out = np.zeros(cap + 16, dtype=np.uint8)
    pos = np.zeros(n, dtype=np.uint64)
    total = C.c_uint64(0)
    idx = None if tile_index is None else np.ascontiguousarray(tile_index, dtype=np.int32)
    check(L.g4_pack_tile_records(context._h, G4_MEM_HOST, a.ctypes.data, off.ctypes.data, ln.ctypes.data,
                                 None if idx is None else idx.ctypes.data, int(first_tile_index), n, 1 if checksum else 0, int(base_pos),
                                 out.ctypes.data, cap, pos.ctypes.data, C.byref(total)), "g4_pack_tile_records")
    return out[:total.value].tobytes(), pos


def unpack_tile_records(context, image, content_pos, checksum):
    a = _u8(image)
    pos = np.ascontiguousarray(content_pos, dtype=np.uint64)
    n = len(pos)
    payload = np.zeros(n, dtype=np.uint64)
    lens = np.zeros(n, dtype=np.uint32)
    status = np.zeros(n, dtype=np.int32)
    check(_lib.lib().g4_unpack_tile_records(context._h, G4_MEM_HOST, a.ctypes.data, a.size, pos.ctypes.data, n, 1 if checksum else 0,
                                            payload.ctypes.data, lens.ctypes.data, status.ctypes.data), "g4_unpack_tile_records")
    return payload, lens, status


def pack_tile_records_device(context, batch, base_pos, checksum, first_tile_index=0):
    """Device-resident form: `batch` is the TileBatch of CodecMaster.encodeTiles on a CUDA tensor.  Returns (records,
    content positions) as torch CUDA tensors; nothing crosses PCIe except the 8-byte total."""
    import torch

    from ._lib import G4_MEM_DEVICE

    L = _lib.lib()
    n = int(batch.lens.numel())
    cap = int(L.g4_tile_records_bound(n, int(batch.total_bytes)))
    dev = batch.arena.device
    out = torch.empty(cap + 16, dtype=torch.uint8, device=dev)
    pos = torch.empty(n, dtype=torch.int64, device=dev)
    total = C.c_uint64(0)
    check(L.g4_pack_tile_records(context._h, G4_MEM_DEVICE, batch.arena.data_ptr(), batch.offsets.data_ptr(), batch.lens.data_ptr(), None,
                                 int(first_tile_index), n, 1 if checksum else 0, int(base_pos), out.data_ptr(), cap, pos.data_ptr(),
                                 C.byref(total)), "g4_pack_tile_records")
    return out[:total.value], pos


def unpack_tile_records_device(context, image, content_pos, checksum):
    """Device-resident form of unpack_tile_records: image and content_pos are torch CUDA tensors (uint8 / int64)."""
    import torch

    from ._lib import G4_MEM_DEVICE

    n = int(content_pos.numel())
    dev = image.device
    payload = torch.empty(n, dtype=torch.int64, device=dev)
    lens = torch.empty(n, dtype=torch.int32, device=dev)
    status = torch.empty(n, dtype=torch.int32, device=dev)
    check(_lib.lib().g4_unpack_tile_records(context._h, G4_MEM_DEVICE, image.data_ptr(), int(image.numel()), content_pos.data_ptr(), n,
                                            1 if checksum else 0, payload.data_ptr(), lens.data_ptr(), status.data_ptr()),
          "g4_unpack_tile_records")
    return payload, lens, status


# ---- writing ---------------------------------------------------------------------------------------------------------
def frame_record(type_code, content, size=None):
    """[size][type][0,0,0] content, zero padding, checksum field left zero (fileSpaceInitRecord / FinishRecord)."""
    if size is None:
        size = multiple_of_8(len(content) + 12)
    return struct.pack("<iB3x", size, type_code) + bytes(content) + bytes(size - 8 - len(content))


def metadata_content(name, record_id, type_code, content, description=""):
    """GvrsMetadata.write (:217-234)."""
    return _utf(name) + struct.pack("<iB3xi", record_id, type_code, len(content)) + bytes(content) + _utf(description)


def codec_metadata(codecs, java_classes=None):
    """The two records GvrsFile writes for a compressed file (GvrsFile.java:301-319): the Java class names per codec id
    ("GvrsJavaCodecs") and the id list ("GvrsCompressionCodecs")."""
    std = {
        "GvrsHuffman": ("org.gridfour.compress.CodecHuffman", "org.gridfour.compress.CodecHuffman"),
        "GvrsDeflate": ("org.gridfour.compress.CodecDeflate", "org.gridfour.compress.CodecDeflate"),
        "GvrsFloat": ("org.gridfour.compress.CodecFloat", "org.gridfour.compress.CodecFloat"),
        "GvrsCanonicalHuffman": ("org.gridfour.compress.canonicalHuffman.CodecCanonHuffman",) * 2,
        "LSOP12": ("org.gridfour.lsop.LsEncoder12", "org.gridfour.lsop.LsDecoder12"),
    }
    std.update(java_classes or {})
    java = "".join("%s,%s,%s\n" % (c, std[c][0], std[c][1]) for c in codecs).encode("utf-8")
    ids = "|".join(codecs).encode("utf-8")
    return [("GvrsJavaCodecs", 0, METADATA_TYPE_STRING, struct.pack("<i", len(java)) + java, "Class paths for Java compressors"),
            ("GvrsCompressionCodecs", 0, METADATA_TYPE_STRING, struct.pack("<i", len(ids)) + ids, "Compession codecs")]


def tile_directory_content(spec, positions, extended=False):
    """RecordManager.writeTileDirectory (:864-883) + TileDirectory.writeTilePositions (:266-278): the bounding box of the
    populated tiles, content position / 8 per cell."""
    cells = {}
    for t, p in positions.items():
        if p:
            cells[divmod(int(t), spec.tiles_across)] = int(p)
    out = bytearray([0, 1 if extended else 0, 0, 0, 0, 0, 0, 0])
    if not cells:
        return bytes(out + struct.pack("<iiii", 0, 0, 0, 0))
    rows, cols = [r for r, _ in cells], [c for _, c in cells]
    row0, col0, n_rows, n_cols = min(rows), min(cols), max(rows) - min(rows) + 1, max(cols) - min(cols) + 1
    out += struct.pack("<iiii", row0, col0, n_rows, n_cols)
    for i in range(n_rows):
        for j in range(n_cols):
            p = cells.get((row0 + i, col0 + j), 0)
            out += struct.pack("<q", p) if extended else struct.pack("<I", (p // 8) & 0xffffffff)
    return bytes(out)


def metadata_directory_content(entries):
    """RecordManager.writeMetadataDirectory (:991-1017); entries = (content position, name, record id, type code), in
    order of file position."""
    out = bytearray(struct.pack("<i", len(entries)))
    for pos, name, record_id, type_code in entries:
        out += struct.pack("<q", pos) + _utf(name) + struct.pack("<iB", record_id, type_code)
    return bytes(out)


class GvrsWriter:
    """Assembles a GVRS file image: header, records in the order they are added, directories, checksums.  Mirrors the
    sequence GvrsFile(File, spec) ... close() produces (GvrsFile.java:241-300, 560-623)."""

    def __init__(self, spec, uuid=None, time_modified=0, version=(1, 4)):
        self.spec, self.version = spec, version
        self.uuid = uuid if uuid is not None else bytes(16)
        self.time_modified = time_modified
        head = bytearray(b"gvrs raster".ljust(12, b"\0") + bytes([version[0], version[1], 0, 0]))
        head += struct.pack("<iB3x", 0, RECORD_HEADER) + self.uuid + struct.pack("<qq", time_modified, 0)
        head += struct.pack("<qq", 0, 0) + struct.pack("<H6x", 1) + struct.pack("<q", 0) + bytes(16)
        head += spec.serialize(len(head)) + bytes(8)
        content_pos = (len(head) + 4 + 7) & ~7
        head += bytes(content_pos - len(head))
        struct.pack_into("<i", head, FILEPOS_HEADER_RECORD, content_pos - FILEPOS_HEADER_RECORD)
        self.buf = head
        self.header_size = content_pos - FILEPOS_HEADER_RECORD
        self.record_spans = [(FILEPOS_HEADER_RECORD, self.header_size)]
        self.metadata_entries = []
        self.tile_positions = {}

    @property
    def file_pos(self):
        return len(self.buf)

    def add_record(self, type_code, content, size=None):
        pos = len(self.buf)
        rec = frame_record(type_code, content, size)
        self.buf += rec
        self.record_spans.append((pos, len(rec)))
        return pos + 8

    def add_metadata(self, name, record_id, type_code, content, description="", size=None):
        pos = self.add_record(RECORD_METADATA, metadata_content(name, record_id, type_code, content, description), size)
        self.metadata_entries.append((pos, name, record_id, type_code))
        return pos

    def add_tile_records(self, context, arena, offsets, lens, tile_index=None, first_tile_index=0):
        """Tile records of a batch, framed (and checksummed) on the GPU."""
        base = len(self.buf)
        recs, pos = pack_tile_records(context, arena, offsets, lens, base, self.spec.checksum, tile_index, first_tile_index)
        self.buf += recs
        idx = range(first_tile_index, first_tile_index + len(pos)) if tile_index is None else tile_index
        for t, p in zip(idx, pos):
            self.tile_positions[int(t)] = int(p)
        return pos

    def add_raster(self, master, grid):
        """A whole one-element raster (numpy, n_rows x n_cols): the cells of the edge tiles that lie outside the raster hold
        the element's fill value, as a RasterTile that was never written there does (TileElementInt/Float/Short
        constructors fill the tile); encodeTiles on the GPU, then the tile records.  Returns the TileBatch."""
        spec = self.spec
        if len(spec.elements) != 1 or grid.shape != (spec.n_rows, spec.n_cols):
            raise ValueError("add_raster: a one-element raster of the specified dimensions")
        e = spec.elements[0]
        dtype = {ELEM_INTEGER: np.int32, ELEM_INT_CODED_FLOAT: np.int32, ELEM_FLOAT: np.float32, ELEM_SHORT: np.int16}[e.type_code]
        full = np.full((spec.tiles_down * spec.tile_rows, spec.tiles_across * spec.tile_cols), e.fill_value, dtype=dtype)
        full[:spec.n_rows, :spec.n_cols] = grid
        batch = master.encodeTiles(full, spec.tile_rows, spec.tile_cols,
                                   fillValue=e.fill_value if e.type_code == ELEM_SHORT else None)
        self.add_tile_records(master._context(), batch.arena, batch.offsets, batch.lens)
        return batch

    def add_tile_record_host(self, tile_index, element_payloads, size=None):
        """One tile record with any number of elements: [tileIndex] then [len][bytes] per element (multi-element files)."""
        content = struct.pack("<i", tile_index) + b"".join(struct.pack("<i", len(p)) + bytes(p) for p in element_payloads)
        pos = self.add_record(RECORD_TILE, content, size)
        self.tile_positions[int(tile_index)] = pos
        return pos

    def finish(self, context, extended_directory=False, directory_order=("metadata", "tile")):
        """close(): metadata directory, tile directory, header slots, then every checksum -- all CRCs in one GPU call.
        (context=None leaves the checksum fields of the host-framed records zero: layout-only use in the CPU tests.)"""
        md_pos = td_pos = 0
        for which in directory_order:
            if which == "metadata" and self.metadata_entries:
                md_pos = self.add_record(RECORD_METADATA_DIR, metadata_directory_content(sorted(self.metadata_entries)))
            elif which == "tile":
                td_pos = self.add_record(RECORD_TILE_DIR, tile_directory_content(self.spec, self.tile_positions, extended_directory))
        struct.pack_into("<qq", self.buf, FILEPOS_FREESPACE_DIR, 0, md_pos)
        struct.pack_into("<q", self.buf, FILEPOS_TILE_DIR, td_pos)
        if self.spec.checksum and context is not None:
            off = np.array([p for p, _ in self.record_spans], dtype=np.uint64)
            size = np.array([s - 4 for _, s in self.record_spans], dtype=np.uint32)
            crc = crc32c_ranges(context, self.buf, off, size)
            for (p, s), c in zip(self.record_spans, crc):
                struct.pack_into("<I", self.buf, p + s - 4, int(c))
        return bytes(self.buf)
