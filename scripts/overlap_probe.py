import time, torch
x = torch.empty(40_000_000, dtype=torch.int32, pin_memory=True)
d = torch.empty_like(x, device="cuda")
a = torch.randn(8192, 8192, device="cuda", dtype=torch.float32)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def work():
    with torch.cuda.stream(s2):
        for _ in range(3): torch.mm(a, a)
def copy():
    with torch.cuda.stream(s1):
        d.copy_(x, non_blocking=True)
for name, f in (("copy", lambda: copy()), ("work", lambda: work()), ("both", lambda: (copy(), work()))):
    for _ in range(3):
        torch.cuda.synchronize(); t = time.perf_counter(); f(); torch.cuda.synchronize(); dt = (time.perf_counter() - t) * 1e3
    print(name, round(dt, 3), "ms")
