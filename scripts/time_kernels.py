"""Per-kernel CUDA-event timings of one fused ArcFace step at the bench shape (B=1024, C=2M), a few repetitions."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import face_recognition_models_b200 as pkg
from face_recognition_models_b200 import _lib as L
B, Cn = int(os.environ.get("B", 1024)), int(os.environ.get("C", 2_000_000))
FAM = os.environ.get("FAM", "arcface")
head = {"arcface": lambda: pkg.ArcFace(512, Cn, s=64.0, m=0.5, easy_margin=False),
        "cosface": lambda: pkg.CosFace(512, Cn, s=64.0, m=0.35),
        "magface": lambda: pkg.MagFace(512, Cn),
        "adaface": lambda: pkg.AdaFace(512, Cn),
        "elastic_arc": lambda: pkg.ElasticArcFace(512, Cn),
        "elastic_cos": lambda: pkg.ElasticCosFace(512, Cn),
        "mv_am": lambda: pkg.MV_Softmax(512, Cn, margin_type="am"),
        "curricularface": lambda: pkg.CurricularFace(512, Cn),
        "sphereface": lambda: pkg.SphereFace(512, Cn, m=2)}[FAM]().cuda()
g = torch.Generator(device="cuda").manual_seed(4)
with torch.no_grad():
    head._param().normal_(0, 0.01, generator=g)
x = torch.randn(B, 512, device="cuda", generator=g) * 3.0
y = torch.randint(0, Cn, (B,), device="cuda", generator=g)
def step():
    xg = x.detach().requires_grad_(True); head._param().grad = None
    out = head.fused_loss(xg, y); out.loss.backward(); return out
def measure(tag):
    for _ in range(5): step()
    torch.cuda.synchronize()
    L.PROFILE = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): step()
    e1.record(); torch.cuda.synchronize()
    prof, L.PROFILE = L.PROFILE, None
    agg = {}
    for name, a, b, n in prof:
        if n: agg.setdefault(name, []).append(a.elapsed_time(b))
    print(f"{tag:28s} step_ms %.3f" % (e0.elapsed_time(e1) / 20), " ".join(f"{k[3:]}={sum(v)/len(v):.3f}" for k, v in agg.items() if sum(v)/len(v) > 0.05), flush=True)

# settings: "ENV=VAL,ENV=VAL;..." (MH_BACKWARD sets head.backward_mode); each measured twice, interleaved
settings = [dict(kv.split("=") for kv in grp.split(",") if kv) for grp in os.environ.get("SETTINGS", "").split(";")]
for rep in range(int(os.environ.get("REPS", 2))):
    for st in settings:
        for k in [k for k in os.environ if k.startswith("MH_") and k not in ("MH_LIB", "MH_DXDW_FRAC", "MH_PROG_AHEAD", "MH_PW_PAIRS")]:
            os.environ.pop(k, None)      # experiment toggles are per setting
        head.backward_mode = "auto"
        for k, v in st.items():
            if k == "MH_BACKWARD": head.backward_mode = v
            else: os.environ[k] = v
        measure(",".join(f"{k}={v}" for k, v in st.items()) or "default")
