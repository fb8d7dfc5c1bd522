"""not-gpu: host-side logic of the N>1 path (tile-row sharding, length gather for the file layout) with a
world_size-2 gloo group on CPU.  The data path itself has no collective."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from gridfour_b200.sharding import gather_layout, record_offsets, shard_tile_rows, tile_content, tile_record_is_compressed


def test_shard_tile_rows_partitions_every_row_once():
    for total, world in ((240, 8), (240, 1), (48, 5), (7, 8), (30, 4)):
        seen = []
        for r in range(world):
            first, n = shard_tile_rows(total, world, r)
            seen += list(range(first, first + n))
        assert seen == list(range(total))
    assert shard_tile_rows(240, 8, 3) == (90, 30)  # config 3: 30 tile rows per GPU


def test_record_offsets_round_to_8_bytes():
    off, total = record_offsets([10, 8, 1, 172800])
    assert off.tolist() == [0, 16, 24, 32] and total == 32 + 172800


def test_tile_record_rule_matches_record_manager():
    """RecordManager.writeTile (:403-461): compressed form only if 4 + sum(4 + len) < 4 + 4E + standard size."""
    n = 90 * 120 * 4
    assert tile_record_is_compressed([n - 1], n) == (True, 4 + 4 + n - 1)
    assert tile_record_is_compressed([n], n) == (False, 4 + 4 + n)  # raw element: same size, not "compressed"
    assert tile_record_is_compressed([100, n], 2 * n) == (True, 4 + 8 + 100 + n)
    assert tile_content([b"\x01\x02\x03", b""]) == b"\x03\x00\x00\x00\x01\x02\x03\x00\x00\x00\x00"


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total_tile_rows, tiles_across = 5, 3  # uneven split: rank 0 gets 3 tile rows, rank 1 gets 2
        first, n = shard_tile_rows(total_tile_rows, world, rank)
        rng = np.random.default_rng(1234)
        all_lens = rng.integers(6, 5000, total_tile_rows * tiles_across)
        mine = all_lens[first * tiles_across:(first + n) * tiles_across]
        lens, off, base, total = gather_layout(mine)
        q.put((rank, lens.tolist(), off.tolist(), base, total, first, n))
    finally:
        dist.destroy_process_group()


def test_gather_layout_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    rng = np.random.default_rng(1234)
    all_lens = rng.integers(6, 5000, 15)
    exp_off, exp_total = record_offsets(all_lens)
    for rank, lens, off, base, total, first, n in res:
        assert lens == all_lens.tolist()
        assert off == exp_off.tolist() and total == exp_total
        assert base == int(exp_off[first * 3])
    assert (res[0][5], res[0][6]) == (0, 3) and (res[1][5], res[1][6]) == (3, 2)
