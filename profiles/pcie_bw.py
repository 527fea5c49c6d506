"""PCIe ceiling for the end-to-end number: pinned host <-> device copies, one direction and both."""
import torch, time
n = 512 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / reps
    return n / dt / 1e9
for _ in range(2): run(True, True)
print("H2D only  %.1f GB/s" % run(True, False))
print("D2H only  %.1f GB/s" % run(False, True))
print("both      %.1f GB/s each direction" % run(True, True))
