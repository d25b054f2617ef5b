"""PCIe microbenchmark for the literal-mode copy plan: contiguous cudaMemcpyAsync against cudaMemcpy2DAsync with narrow rows
(eos_vars rows 1-3 of 7: width 24 B at pitch 56 B; the h column of xyzh: width 8 B at pitch 32 B), both directions."""
import sys
import time

import torch
from cuda.bindings import runtime as rt

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2097152
dev = torch.empty(n * 7, dtype=torch.float64, device="cuda")
host = torch.empty(n * 7, dtype=torch.float64).pin_memory()
st = torch.cuda.current_stream().cuda_stream
D2H, H2D = rt.cudaMemcpyKind.cudaMemcpyDeviceToHost, rt.cudaMemcpyKind.cudaMemcpyHostToDevice


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best


for label, width, dpitch, spitch in (("contiguous 56B/row", 56, 56, 56), ("eos rows 24B of 56B (host strided)", 24, 56, 24), ("eos rows 24B of 56B (both strided)", 24, 56, 56),
                                     ("h column 8B of 32B (host strided)", 8, 32, 8), ("xyz 24B of 32B (both strided)", 24, 32, 32)):
    t = timed(lambda: rt.cudaMemcpy2DAsync(host.data_ptr(), dpitch, dev.data_ptr(), spitch, width, n, D2H, st))
    t2 = timed(lambda: rt.cudaMemcpy2DAsync(dev.data_ptr(), spitch, host.data_ptr(), dpitch, width, n, H2D, st))
    print(f"{label:40s} D2H {t*1e3:7.3f} ms ({width*n/t/1e9:6.1f} GB/s payload)   H2D {t2*1e3:7.3f} ms ({width*n/t2/1e9:6.1f} GB/s payload)")
