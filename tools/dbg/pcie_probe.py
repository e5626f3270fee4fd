import torch, time
n = 1920*1080*3
h_in = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(4)]
h_out = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(4)]
d_a = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(4)]
d_b = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(4)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
print("asyncEngineCount", torch.cuda.get_device_properties(0).multi_processor_count, getattr(torch.cuda.get_device_properties(0), "async_engine_count", "?"))
def bench(fn, iters=50):
    torch.cuda.synchronize(); t=time.perf_counter()
    for i in range(iters): fn(i)
    torch.cuda.synchronize(); return (time.perf_counter()-t)/iters*1e6
def h2d(i):
    with torch.cuda.stream(s1): d_a[i%4].copy_(h_in[i%4], non_blocking=True)
def d2h(i):
    with torch.cuda.stream(s2): h_out[i%4].copy_(d_b[i%4], non_blocking=True)
def both(i): h2d(i); d2h(i)
for name, fn in (("H2D", h2d), ("D2H", d2h), ("both concurrently", both)):
    us = bench(fn); print(f"{name}: {us:.1f} us per frame-sized copy -> {n/us/1e3:.1f} GB/s per direction")
