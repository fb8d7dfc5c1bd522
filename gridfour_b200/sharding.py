"""Tile-row sharding of a raster over the GPUs of one box.

A tile's packing depends only on its own cells (compress/PredictorModelDifferencing.java:119 seeds from the tile's
first cell; LSOP coefficients are per tile, lsop/LsOptimalPredictor12.java:212), so GPU g simply owns the contiguous
band of tile rows [g*T/G, (g+1)*T/G) and there is no collective on the data path.  The only exchange is the per-tile
payload lengths that the host needs for the file layout: records are allocated in multiples of 8 bytes
(gvrs/RecordManager.java:137-139,218-219), so global offsets are an exclusive scan of the 8-byte-rounded lengths in
tile order.  (Paths under /root/reference/core/src/main/java/org/gridfour/.)
"""
import numpy as np


def shard_tile_rows(total_tile_rows, world_size, rank):
    """Returns (first_tile_row, n_tile_rows) of `rank`; the remainder is spread over the first ranks."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(total_tile_rows), int(world_size))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def record_offsets(lens):
    """Exclusive scan of 8-byte-rounded payload lengths -> (offsets, total): the arena / file layout of a tile list."""
    lens = np.asarray(lens, dtype=np.uint64)
    padded = (lens + np.uint64(7)) & ~np.uint64(7)
    off = np.zeros(lens.size, np.uint64)
    if lens.size > 1:
        np.cumsum(padded[:-1], out=off[1:])
    return off, int(padded.sum())


def gather_layout(local_lens, group=None):
    """All ranks contribute their band's per-tile lengths (tile order); every rank gets the global lengths, the global
    offsets and its own base offset.  Uses torch.distributed (gloo on CPU tensors, nccl on CUDA tensors); lengths only --
    payload bytes never cross ranks."""
    import torch
    import torch.distributed as dist

    local = torch.as_tensor(np.asarray(local_lens, dtype=np.int64))
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        lens = np.asarray(local_lens, dtype=np.uint64)
        off, total = record_offsets(lens)
        return lens, off, 0, total
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    local = local.to(dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([local.numel()], dtype=torch.int64, device=dev), group=group)
    counts = [int(c.item()) for c in counts]
    width = max(counts)
    padded = torch.zeros(width, dtype=torch.int64, device=dev)
    padded[: local.numel()] = local
    parts = [torch.zeros(width, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    lens = np.concatenate([p[:c].cpu().numpy() for p, c in zip(parts, counts)]).astype(np.uint64)
    off, total = record_offsets(lens)
    rank = dist.get_rank(group)
    first = int(sum(counts[:rank]))
    base = int(off[first]) if first < off.size else total
    return lens, off, base, total


def tile_content(element_payloads):
    """Tile content of a compressed tile record: for every element [len:int32 LE][payload]
    (gvrs/RasterTile.java:234-256, getCompressedPacking).  `element_payloads` = what g4_encode_tiles produced for the
    tile's elements (a payload of exactly the element's standard size is the raw form, gvrs/TileElementInt.java:198-204)."""
    out = bytearray()
    for p in element_payloads:
        out += int(len(p)).to_bytes(4, "little")
        out += bytes(p)
    return bytes(out)


def tile_record_is_compressed(element_lens, standard_tile_bytes):
    """gvrs/RecordManager.java:403-461 (writeTile): the compressed form of a tile (4-byte tile index + [len][payload] per
    element) is stored only if it is strictly smaller than the uncompressed record, 4 + 4*E + standardTileDataSizeInBytes.
    Returns (use_compressed, record_payload_bytes)."""
    e = len(element_lens)
    compressed = 4 + sum(4 + int(n) for n in element_lens)
    payload = 4 + 4 * e + int(standard_tile_bytes)
    return (compressed < payload), (compressed if compressed < payload else payload)
