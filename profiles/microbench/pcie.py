import torch, time
dev = torch.device("cuda:0")
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device=dev)
h2 = torch.empty(n // 4, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n // 4, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
t = run(lambda: d.copy_(h, non_blocking=True)); print("H2D alone  %.1f GB/s" % (n / t / 1e9))
t = run(lambda: h.copy_(d, non_blocking=True)); print("D2H alone  %.1f GB/s" % (n / t / 1e9))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
t = run(both); print("H2D with concurrent D2H (1/4 volume): H2D %.1f GB/s, D2H %.1f GB/s" % (n / t / 1e9, n / 4 / t / 1e9))
for chunk in (8 << 20, 64 << 20, 256 << 20):
    def chunked():
        for o in range(0, n, chunk): d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True)
    t = run(chunked); print("H2D in %d MB chunks %.1f GB/s" % (chunk >> 20, n / t / 1e9))
