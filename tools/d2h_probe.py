import torch, time
n = 8 << 20
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
host = torch.empty(n, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def one(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(s1):
        for _ in range(reps): host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize(); return n * reps / (time.perf_counter() - t0) / 1e9
def two(reps, parts=2):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    h = n // parts
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        for i, s in enumerate(streams):
            with torch.cuda.stream(s): host[i*h:(i+1)*h].copy_(dev[i*h:(i+1)*h], non_blocking=True)
    torch.cuda.synchronize(); return n * reps / (time.perf_counter() - t0) / 1e9
for f in (one, two): f(20)
print("D2H 8 MB, one stream:   %.1f GB/s" % one(400))
print("D2H 2 x 4 MB, 2 streams: %.1f GB/s" % two(400))
print("D2H 4 x 2 MB, 4 streams: %.1f GB/s" % two(400, 4))
print("D2H 8 MB, one stream:   %.1f GB/s" % one(400))
# with concurrent H2D of 1 MB
hin = torch.empty(1 << 20, dtype=torch.uint8).pin_memory(); din = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
s3 = torch.cuda.Stream()
def both(reps, parts):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    h = n // parts
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        with torch.cuda.stream(s3): din.copy_(hin, non_blocking=True)
        for i, s in enumerate(streams):
            with torch.cuda.stream(s): host[i*h:(i+1)*h].copy_(dev[i*h:(i+1)*h], non_blocking=True)
    torch.cuda.synchronize(); return n * reps / (time.perf_counter() - t0) / 1e9
print("with 1 MB H2D alongside: 1 stream %.1f GB/s, 2 streams %.1f GB/s" % (both(400, 1), both(400, 2)))
