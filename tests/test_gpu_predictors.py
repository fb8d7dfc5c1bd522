"""-m gpu: the predictor models on their own (IPredictorModel.java:42-173) -- north_star lists "predictor byte streams are
bit-exact" as its own criterion.  GPU encode / encodeInt against the oracle's restatement of the four Java models
(seed, M32 bytes, residual ints), GPU decode / decodeInt of the ORACLE's streams back to the tile, and the reference's own
predictor unit-test pattern (round trips on its test grids: compress/PredictorModel*Test.java)."""
import numpy as np
import pytest

from gpu_common import NULL, first_diff, parity_grids

pytestmark = pytest.mark.gpu

MODELS = {1: "PredictorModelDifferencing", 2: "PredictorModelLinear", 3: "PredictorModelTriangle"}


@pytest.fixture(scope="module")
def g4():
    import gridfour_b200

    return gridfour_b200


def grids(oracle):
    g = parity_grids(oracle)
    g.pop("two_by_two")
    g["terrain128"] = oracle.terrain_i32(40, 40, 128, 128)
    g["terrain256"] = oracle.terrain_i32(4000, 100, 256, 256)
    return g


@pytest.mark.parametrize("model", [1, 2, 3])
def test_predictor_streams_bit_exact(g4, oracle, model):
    pm = getattr(g4, MODELS[model])()
    assert pm.getPredictorType().getCodeValue() == model and not pm.isNullDataSupported()
    for name, tile in grids(oracle).items():
        nr, nc = tile.shape
        n_want, seed_want, m32_want = oracle.predictor_encode(model, tile)
        buf = np.zeros(6 * tile.size, np.uint8)
        n = pm.encode(nr, nc, tile, buf)
        assert n == n_want, "%s %s: %d vs %d bytes" % (MODELS[model], name, n, n_want)
        assert pm.getSeed() == seed_want
        assert buf[:n].tobytes() == bytes(m32_want), "%s %s: %s" % (MODELS[model], name, first_diff(buf[:n].tobytes(), bytes(m32_want)))
        out = np.zeros((nr, nc), np.int32)
        pm.decode(seed_want, nr, nc, bytes(m32_want), 0, n_want, out)
        assert np.array_equal(out, tile), "%s %s decode" % (MODELS[model], name)
        # int flavour (CodecCanonHuffman's input)
        k_want, seed2, res_want = oracle.predictor_encode_int(model, tile)
        res = np.zeros(tile.size, np.int32)
        k = pm.encodeInt(nr, nc, tile, res)
        assert k == k_want and pm.getSeed() == seed2
        assert np.array_equal(res[:k], np.asarray(res_want)[:k_want]), "%s %s residual ints" % (MODELS[model], name)
        out[:] = 0
        pm.decodeInt(seed2, nr, nc, np.asarray(res_want, dtype=np.int32), 0, k_want, out)
        assert np.array_equal(out, tile)


def test_nulls_model_streams_bit_exact(g4, oracle):
    rng = np.random.default_rng(3)
    pm = g4.PredictorModelDifferencingWithNulls()
    assert pm.isNullDataSupported() and pm.getPredictorType().getCodeValue() == 4
    base = oracle.terrain_i32(10, 10, 64, 96).copy()
    cases = {}
    for frac in (0.01, 0.3, 0.9):
        t = base.copy()
        t[rng.random(t.shape) < frac] = NULL
        cases["frac%.2f" % frac] = t
    t = base.copy(); t[:, 0] = NULL; cases["null first column"] = t
    t = base.copy(); t[0, :] = NULL; t[5:9, :] = NULL; cases["null rows"] = t
    for name, tile in cases.items():
        nr, nc = tile.shape
        n_want, seed_want, m32_want = oracle.predictor_encode(4, tile)
        buf = np.zeros(6 * tile.size, np.uint8)
        n = pm.encode(nr, nc, tile, buf)
        assert (n, pm.getSeed()) == (n_want, seed_want), name
        assert buf[:n].tobytes() == bytes(m32_want), name
        out = np.zeros((nr, nc), np.int32)
        pm.decode(seed_want, nr, nc, bytes(m32_want), 0, n_want, out)
        assert np.array_equal(out, tile), name
    all_null = np.full((8, 8), NULL, np.int32)
    assert pm.encode(8, 8, all_null, np.zeros(6 * 64, np.uint8)) == -1


def test_triangle_declines_a_one_row_tile_and_malformed_m32_is_an_error(g4, oracle):
    tile = oracle.terrain_i32(0, 0, 16, 16)
    pm = g4.PredictorModelTriangle()
    n, seed, m32 = oracle.predictor_encode(3, tile)
    out = np.zeros((16, 16), np.int32)
    with pytest.raises(g4.FormatError):
        pm.decode(seed, 16, 16, bytes(m32)[:n - 3], 0, n - 3, out)  # too few residuals
    bad = bytearray(bytes(m32)[:n])
    bad[-1] = 0x7F  # an introducer with nothing behind it
    with pytest.raises(g4.FormatError):
        pm.decode(seed, 16, 16, bytes(bad), 0, n, out)


def test_predictor_tiles_band_matches_per_tile(g4, oracle):
    """g4_predictor_tiles over a config-1 band on the device: every tile's M32 stream equals the oracle's."""
    import ctypes as C

    import torch

    from gridfour_b200 import _lib

    tr, tc, down, across = 90, 120, 2, 8
    grid = oracle.terrain_i32(0, 0, down * tr, across * tc)
    g = torch.from_numpy(grid).cuda()
    n_tiles, slot = down * across, ((6 * tr * tc + 16 + 15) // 16) * 16
    slots = torch.zeros(n_tiles * slot, dtype=torch.uint8, device="cuda")
    lens = torch.zeros(n_tiles, dtype=torch.int32, device="cuda")
    seeds = torch.zeros(n_tiles, dtype=torch.int32, device="cuda")
    status = torch.zeros(n_tiles, dtype=torch.int32, device="cuda")
    ctx = g4.Context.default()
    band = g4.CodecMaster._band(grid.shape, np.int32, tr, tc)
    torch.cuda.synchronize()
    for model in (1, 2, 3):
        _lib.check(_lib.lib().g4_predictor_tiles(ctx._h, model, 0, 0, C.byref(band), g.data_ptr(), slots.data_ptr(), slot, lens.data_ptr(),
                                                 seeds.data_ptr(), status.data_ptr()))
        ctx.synchronize()
        assert int(status.abs().sum()) == 0
        h, hl, hs = slots.cpu().numpy(), lens.cpu().numpy(), seeds.cpu().numpy()
        for t in range(n_tiles):
            r, c = divmod(t, across)
            n, seed, m32 = oracle.predictor_encode(model, grid[r * tr:(r + 1) * tr, c * tc:(c + 1) * tc])
            assert (hl[t], hs[t]) == (n, seed)
            assert h[t * slot:t * slot + n].tobytes() == bytes(m32)
        out = torch.zeros_like(g)
        _lib.check(_lib.lib().g4_predictor_tiles(ctx._h, model, 0, 1, C.byref(band), out.data_ptr(), slots.data_ptr(), slot, lens.data_ptr(),
                                                 seeds.data_ptr(), status.data_ptr()))
        ctx.synchronize()
        assert int(status.abs().sum()) == 0 and torch.equal(out, g)
