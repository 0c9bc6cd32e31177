import time, torch
dev = torch.device("cuda:0")
h_in = torch.empty(130_000_000, dtype=torch.uint8).pin_memory()
d_in = torch.empty(130_000_000, dtype=torch.uint8, device=dev)
h_out = torch.empty(30_000_000, dtype=torch.uint8).pin_memory()
d_out = torch.empty(30_000_000, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.time() - t0) / reps * 1e3
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both():
    h2d(); d2h()
def h2d_chunks():
    with torch.cuda.stream(s1):
        for i in range(8):
            a, b = i * 16_250_000, (i + 1) * 16_250_000
            d_in[a:b].copy_(h_in[a:b], non_blocking=True)
print("H2D 130MB  %.3f ms  %.1f GB/s" % (t(h2d), 130 / t(h2d)))
print("D2H 30MB   %.3f ms  %.1f GB/s" % (t(d2h), 30 / t(d2h)))
print("both       %.3f ms" % t(both))
print("H2D in 8 chunks %.3f ms" % t(h2d_chunks))
