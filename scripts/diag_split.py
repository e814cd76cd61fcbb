"""Diagnostic: error of one split-precision conv (yb_conv2d precision 3) against float64 for activations whose lo part
would be an fp16 subnormal (|x| < 0.25) vs not.  With fp16 lo planes the tensor core flushes those to zero; with bf16
lo planes both ranges must sit at fp32 level."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from yolact_b200 import _lib
lib = _lib.load()
yc = _lib.YbConfig(); yc.backbone = _lib.YB_BACKBONE_NONE
yc.num_classes, yc.mask_dim, yc.precision = 81, 32, _lib.YB_PREC_F32
yc.nms_top_k, yc.nms_conf_thresh, yc.nms_thresh, yc.max_num_detections = 200, 0.05, 0.5, 100
h = ctypes.c_void_p(); _lib.check(lib.yb_create(ctypes.byref(yc), 0, ctypes.byref(h)), "create")
g = torch.Generator().manual_seed(0)
B, Ci, H, W, Co, k = 2, 256, 35, 35, 256, 3
w = torch.randn(Co, Ci, k, k, generator=g) * (2.0 / (Ci * k * k)) ** 0.5
for name, x in (("relu(N(0,1))", torch.randn(B, Ci, H, W, generator=g).clamp(min=0)),
                ("U[1,2)", torch.rand(B, Ci, H, W, generator=g) + 1.0),
                ("U[0.01,0.02)", torch.rand(B, Ci, H, W, generator=g) * 0.01 + 0.01),
                ("U[1e-4,2e-4)", torch.rand(B, Ci, H, W, generator=g) * 1e-4 + 1e-4)):
    ref = F.conv2d(x.double(), w.double(), padding=1)
    for prec in (3, 1, 0):
        y = torch.empty(B, Co, H, W, device="cuda")
        xd = x.cuda().contiguous()
        _lib.check(lib.yb_conv2d(h, _lib.ptr(xd), ctypes.c_void_p(w.contiguous().data_ptr()), None, None, _lib.ptr(y), B, Ci, H, W,
                                 Co, k, k, 1, 1, 0, prec, 1, None, _lib.current_stream()), "conv2d")
        torch.cuda.synchronize()
        e = (y.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
        print("%-14s precision %d: max err / range = %.3e" % (name, prec, e))
