import torch, time
n = 100*1024*1024//4
h_in = torch.empty(n, dtype=torch.float32).pin_memory(); h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_a = torch.empty(n, dtype=torch.float32, device='cuda'); d_b = torch.empty(n, dtype=torch.float32, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=10):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both(): h2d(); d2h()
gb = n*4/1e9
print("H2D alone %.1f GB/s" % (gb/t(h2d)), "D2H alone %.1f GB/s" % (gb/t(d2h)), "duplex %.1f GB/s each" % (gb/t(both)))
# 2D strided copy like the engine: 24 rows of 512 KB out of a pitch of 4 MB
rows, w, pitch = 24, 131072, 1048576
h2 = torch.empty(rows*pitch, dtype=torch.float32).pin_memory(); d2 = torch.empty(rows*w, dtype=torch.float32, device='cuda')
def h2d_2d():
    with torch.cuda.stream(s1): d2.view(rows, w).copy_(h2.view(rows, pitch)[:, :w], non_blocking=True)
print("2D-strided H2D (24 x 512 KB rows) %.1f GB/s" % (rows*w*4/1e9/t(h2d_2d)))
