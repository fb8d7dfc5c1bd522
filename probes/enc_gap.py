"""Where the encode wall time goes beyond the kernels: event time of encodeTiles per codec list vs the library's own kernel brackets."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gridfour_b200 as g4
sys.path.insert(0, "tests")
dev = torch.device("cuda", 0)
rows, cols = 5400, 86400
grid = torch.empty((rows, cols), dtype=torch.int32, device=dev)
ctx = g4.Context.default(0)
ctx.fill_terrain(grid.data_ptr(), 0, 0, 0, rows, cols)
torch.cuda.synchronize()
ctx.set_timing(True)
for names in (["GvrsHuffman"], ["GvrsDeflate"], ["LSOP12"], ["GvrsHuffman", "GvrsDeflate", "LSOP12"]):
    spec = g4.CodecSpecification(default=False)
    table = {"GvrsHuffman": (g4.CodecHuffman, g4.CodecHuffman), "GvrsDeflate": (g4.CodecDeflate, g4.CodecDeflate), "LSOP12": (g4.LsEncoder12, g4.LsDecoder12)}
    for n in names:
        spec.addCompressionCodec(n, *table[n])
    master = g4.CodecMaster(spec)
    for i in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        b = master.encodeTiles(grid, 180, 240)
        e1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        print(names, "event ms %.1f wall ms %.1f" % (e0.elapsed_time(e1), 1000 * (t1 - t0)), flush=True)
        del b
