"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/g4codec.h
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "g4codec.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(g4_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from gridfour_b200 import _lib

    L = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(L, s), "libg4codec.so does not export %s" % s
    assert sorted(_lib.EXPORTS) == syms
    assert L.g4_abi_version() == 5


def test_codec_names_round_trip():
    from gridfour_b200 import _lib

    L = _lib.lib()
    for i, name in enumerate([b"GvrsHuffman", b"GvrsDeflate", b"GvrsFloat", b"GvrsCanonicalHuffman", b"LSOP12", b"LSOP08"]):
        assert L.g4_codec_id_from_name(name) == i
        assert L.g4_codec_name(i) == name
    assert L.g4_codec_id_from_name(b"nope") == -1
    # LSOP08 is the legacy decode-only codec (lsop/LsCodecUtility.java:73): kernels for decode (direction 0), none for encode
    assert L.g4_codec_supported(5, 0) == 1 and L.g4_codec_supported(5, 1) == 0


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import gridfour_b200 as g

    with pytest.raises(g.G4Error) as e:
        g.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_codec_specification_mirrors_reference_rules():
    """GvrsFileSpecification.addCompressionCodec: same id replaces and moves to the end (java :1590,1628-1630)."""
    import gridfour_b200 as g

    spec = g.CodecSpecification()
    assert [c[0] for c in spec.getCompressionCodecs()] == ["GvrsHuffman", "GvrsDeflate", "GvrsFloat"]
    spec.addCompressionCodec("GvrsHuffman", g.CodecHuffman)
    assert [c[0] for c in spec.getCompressionCodecs()] == ["GvrsDeflate", "GvrsFloat", "GvrsHuffman"]
    spec.addCompressionCodec("LSOP12", g.LsEncoder12, g.LsDecoder12)
    cl = spec.native_list()
    assert cl.n_codecs == 4 and list(cl.codec_ids[:4]) == [1, 2, 0, 4]
    with pytest.raises(ValueError):
        spec.addCompressionCodec("bad id!", g.CodecHuffman)
    with pytest.raises(ValueError):
        spec.addCompressionCodec("x", int)
    spec.removeAllCompressionCodecs()
    assert spec.native_list().n_codecs == 0
