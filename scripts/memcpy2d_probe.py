"""Is a strided D2H copy of the output span of struct part faster than the whole AoS?"""
import time, torch
from cuda import cudart
n, pitch = 2097152, 128
dev = torch.empty(n * pitch, dtype=torch.uint8, device="cuda")
host = torch.empty(n * pitch, dtype=torch.uint8).pin_memory()
s = torch.cuda.current_stream().cuda_stream
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
full = lambda: cudart.cudaMemcpyAsync(host.data_ptr(), dev.data_ptr(), n * pitch, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost, s)
print("full D2H 268 MB: %.3f ms" % timed(full))
for off, width in ((48, 72), (52, 65), (48, 80), (16, 104), (64, 64)):
    f = lambda: cudart.cudaMemcpy2DAsync(host.data_ptr() + off, pitch, dev.data_ptr() + off, pitch, width, n, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost, s)
    print("2D D2H off %d width %d: %.3f ms" % (off, width, timed(f)))
    f = lambda: cudart.cudaMemcpy2DAsync(dev.data_ptr() + off, pitch, host.data_ptr() + off, pitch, width, n, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, s)
    print("2D H2D off %d width %d: %.3f ms" % (off, width, timed(f)))
h2d = lambda: cudart.cudaMemcpyAsync(dev.data_ptr(), host.data_ptr(), n * pitch, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, s)
print("full H2D 268 MB: %.3f ms" % timed(h2d))
