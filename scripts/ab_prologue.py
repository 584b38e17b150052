"""A/B of mh_prologue_w (and mh_sgd_step_w when exported) between builds of the library on one box:
  python scripts/ab_prologue.py libA.so libB.so ...   (env C, LAYOUT=0|1)"""
import ctypes as C
import os
import sys

import torch

Cn = int(os.environ.get("C", 2_000_000))
layout = int(os.environ.get("LAYOUT", 1))
C_pad = (Cn + 255) // 256 * 256
shape = (Cn, 512) if layout == 0 else (512, Cn)
W = torch.randn(shape, device="cuda") * 0.01
G = torch.randn(shape, device="cuda") * 0.001
M = torch.zeros(shape, device="cuda")
w_hat = torch.empty(C_pad, 512, dtype=torch.bfloat16, device="cuda")
inv = torch.empty(Cn, device="cuda")
vp = C.c_void_p
st = vp(torch.cuda.current_stream().cuda_stream)
libs = [(p, C.CDLL(os.path.abspath(p))) for p in sys.argv[1:]]
ref = None
for rnd in range(3):
    for path, lib in libs:
        def pro():
            rc = lib.mh_prologue_w(vp(W.data_ptr()), C.c_int(layout), C.c_int64(Cn), C.c_int64(shape[1]), vp(w_hat.data_ptr()),
                                   C.c_int64(C_pad), vp(0), vp(inv.data_ptr()), st)
            assert rc == 0

        def sgd():
            rc = lib.mh_sgd_step_w(vp(W.data_ptr()), C.c_int(layout), C.c_int64(Cn), C.c_int64(shape[1]), vp(G.data_ptr()),
                                   vp(M.data_ptr()), C.c_float(1e-4), C.c_float(0.9), C.c_float(5e-4), vp(0), vp(0),
                                   vp(w_hat.data_ptr()), C.c_int64(C_pad), vp(inv.data_ptr()), st)
            assert rc == 0
        fns = [("prologue_w", pro)] + ([("sgd_step_w", sgd)] if hasattr(lib, "mh_sgd_step_w") else [])
        for name, fn in fns:
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            if name == "prologue_w":
                if ref is None:
                    ref = (w_hat.clone(), inv.clone())
                same = torch.equal(ref[0], w_hat) and torch.equal(ref[1], inv)
            else:
                same = "-"
            nbytes = Cn * 512 * (6 if name == "prologue_w" else 22)
            print(f"round {rnd} {os.path.basename(path):32s} {name:12s} {ms:7.3f} ms  {nbytes / ms / 1e6:7.1f} GB/s  same_bits={same}")
    ref = None if rnd < 2 else ref
