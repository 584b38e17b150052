"""Host-side profile of the public-API step at a launch-bound shape (BASELINE config 2: B=512, C=10,575): where the
Python time of `fused_loss(...).loss.backward()` goes.  cProfile, top entries by cumulative and by own time."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import face_recognition_models_b200 as pkg

B, Cn = int(os.environ.get("B", 512)), int(os.environ.get("C", 10575))
head = pkg.CosFace(512, Cn, s=64.0, m=0.35).cuda()
x = torch.randn(B, 512, device="cuda")
y = torch.randint(0, Cn, (B,), device="cuda")


def step():
    xg = x.detach().requires_grad_(True)
    head.kernel.grad = None
    out = head.fused_loss(xg, y)
    out.loss.backward()


for _ in range(50):
    step()
torch.cuda.synchronize()
N = 2000
t0 = time.perf_counter()
for _ in range(N):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host-issue time per step {1e6 * (t1 - t0) / N:.1f} us; incl. drain {1e6 * (t2 - t0) / N:.1f} us")
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
st.sort_stats("tottime").print_stats(18)
