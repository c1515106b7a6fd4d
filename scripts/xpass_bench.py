"""Micro-benchmark of trvb_box_fields_fused (2-D D2Z, k_xpass_fused, 2-D Z2D) on a random REAL
mesh through the C ABI: per-step CUDA-event times for the kernel variants
(TRV_XPASS_VARIANT), and equality of their outputs.  usage: xpass_bench.py [ngrid] [nsub]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from triumvirate_b200 import _lib  # noqa: E402

ng = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 144
variants = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 2, 3]
tb = _lib.trvb()
dev = torch.device("cuda", 0)


def chk(st):
    if st != 0:
        raise RuntimeError(tb.trvb_last_error().decode())


class Mesh(C.Structure):
    _fields_ = [("data", C.c_void_p), ("layout", C.c_int), ("k0_add", C.c_double)]


ctx, sub = C.c_void_p(), C.c_void_p()
chk(tb.trvb_ctx_create(C.byref(ctx), 0, (C.c_int * 3)(ng, ng, ng), (C.c_double * 3)(1000., 1000., 1000.), 4))
chk(tb.trvb_subgrid_create(ctx, C.byref(sub), (C.c_int * 3)(ns, ns, ns)))
gen = torch.Generator(device=dev).manual_seed(5)
x = torch.rand(ng, ng, ng, dtype=torch.float64, device=dev, generator=gen)
xi = torch.empty(ng, ng, ng, dtype=torch.float64, device=dev)
lowk = torch.empty(ns, ns, ns // 2 + 1, 2, dtype=torch.float64, device=dev)
S = (C.c_double * 2)(1.e7, 0.)
os.environ["TRV_XPASS_TRACE"] = "1"
ref = None
for v in variants:
    os.environ["TRV_XPASS_VARIANT"] = str(v)
    ms = []
    for r in range(8):
        chk(tb.trvb_box_fields_fused(ctx, sub, Mesh(x.data_ptr(), 0, 0.), C.c_double(-1.e7), C.c_double(0.), S,
                                     Mesh(lowk.data_ptr(), 2, 0.), Mesh(xi.data_ptr(), 0, 0.)))
        buf = (C.c_double * 3)()
        tb.trvb_box_fields_fused_last_ms(buf)
        ms.append(buf[:])
    tb.trvb_ctx_forget_lowk(ctx)
    med = np.median(np.array(ms[2:]), axis=0)
    cur = (xi.clone(), lowk.clone())
    line = f"variant {v}: d2z_2d {med[0]:.3f} ms, k_xpass_fused {med[1]:.3f} ms, z2d_2d {med[2]:.3f} ms"
    if ref is None:
        ref = cur
    else:
        e1 = float((cur[0] - ref[0]).abs().max() / ref[0].abs().max())
        e2 = float((cur[1] - ref[1]).abs().max() / ref[1].abs().max())
        line += f"; vs variant {variants[0]}: xi {e1:.1e} lowk {e2:.1e}"
    print(line, flush=True)
