"""Shared helpers for the -m gpu parity tests (CUDA path vs CPU oracle on the same seeded inputs)."""
import numpy as np

NULL = -(2 ** 31)


def parity_grids(oracle, small=True):
    rng = np.random.default_rng(7)
    r, c = np.mgrid[0:10, 0:10]
    g = {}
    g["ref10x10"] = (r * 10 + c).astype(np.int32)
    g["terrain45x60"] = oracle.terrain_i32(1000, 2000, 45, 60)
    g["terrain90x120"] = oracle.terrain_i32(0, 0, 90, 120)
    g["terrain180x240"] = oracle.terrain_i32(7000, 3000, 180, 240)
    g["terrain61x77"] = oracle.terrain_i32(123, 457, 61, 77)  # odd sizes: unaligned rows
    g["const"] = np.full((12, 17), 42, np.int32)
    g["noise"] = rng.integers(-(2 ** 31), 2 ** 31, (20, 30), dtype=np.int64).astype(np.int32)
    g["checker"] = np.where((r + c) % 2 == 0, 2 ** 31 - 1, -(2 ** 31) + 1).astype(np.int32)
    g["ramp"] = (r * 1000 - c * 77).astype(np.int32)
    g["smallnoise"] = rng.integers(-3, 4, (33, 47)).astype(np.int32)
    g["wide"] = rng.integers(-40000, 40000, (31, 29)).astype(np.int32)
    g["two_by_two"] = np.array([[1, 2], [3, 5]], np.int32)
    g["steps"] = (np.add.outer(np.arange(64) // 8, np.arange(96) // 8) * 300).astype(np.int32)
    g["skewed"] = np.where(rng.random((70, 70)) < 0.97, 0, rng.integers(-10 ** 6, 10 ** 6, (70, 70))).cumsum(axis=1).astype(np.int32)
    return g


def first_diff(a, b):
    a = np.frombuffer(a, np.uint8) if isinstance(a, (bytes, bytearray)) else np.asarray(a).ravel()
    b = np.frombuffer(b, np.uint8) if isinstance(b, (bytes, bytearray)) else np.asarray(b).ravel()
    n = min(a.size, b.size)
    d = np.nonzero(a[:n] != b[:n])[0]
    return "sizes %d vs %d, first diff at %s" % (a.size, b.size, d[0] if d.size else None)
