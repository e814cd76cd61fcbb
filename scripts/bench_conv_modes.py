"""Kernel time of single conv layers under explicit plan overrides (yb_conv2d, iters=30): which of plain / pair /
stream-K / A-stationary wins per layer shape, outside the autotuner.  Usage: python scripts/bench_conv_modes.py
HISTORICAL: written for the build that still had the A-stationary variant (YB_CONV2D_SK=2); it lost everywhere
(profiles/conv_modes_r02.md) and was removed, so on the current library the "astat" rows repeat the stream-K ones."""
import ctypes, itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolact_b200 import _lib
lib = _lib.load()
yc = _lib.YbConfig(); yc.backbone = _lib.YB_BACKBONE_NONE
yc.num_classes, yc.mask_dim, yc.precision = 81, 32, _lib.YB_PREC_F32
yc.nms_top_k, yc.nms_conf_thresh, yc.nms_thresh, yc.max_num_detections = 200, 0.05, 0.5, 100
h = ctypes.c_void_p(); _lib.check(lib.yb_create(ctypes.byref(yc), 0, ctypes.byref(h)), "create")
SHAPES = [("s3 conv3 256->1024 1x1 +res", 8, 256, 35, 1024, 1, True), ("s3 conv1 1024->256 1x1", 8, 1024, 35, 256, 1, False),
          ("s3 conv2 256->256 3x3", 8, 256, 35, 256, 3, False), ("s2 conv3 128->512 1x1 +res", 8, 128, 69, 512, 1, True),
          ("s1 conv3 64->256 1x1 +res", 8, 64, 138, 256, 1, True)]
MODES = [("plain", {}), ("pair", {"YB_CONV2D_PAIR": "1"}), ("sk", {"YB_CONV2D_SK": "1"}), ("sk+pair", {"YB_CONV2D_SK": "1", "YB_CONV2D_PAIR": "1"}),
         ("astat", {"YB_CONV2D_SK": "2"}), ("astat+pair", {"YB_CONV2D_SK": "2", "YB_CONV2D_PAIR": "1"})]
KEYS = ["YB_CONV2D_PAIR", "YB_CONV2D_SK", "YB_CONV2D_BN", "YB_CONV2D_EPI"]
g = torch.Generator().manual_seed(0)
for prec in (3, 1):
    print("\n## precision %d (%s)\n" % (prec, "split f16x3" if prec == 3 else "f16tc"))
    print("| layer | mode | BN=64 | BN=128 | BN=256 | BN=128 epi2 |\n|---|---|---:|---:|---:|---:|")
    for name, B, Ci, HW, Co, k, res in SHAPES:
        x = torch.randn(B, Ci, HW, HW, generator=g).cuda()
        w = (torch.randn(Co, Ci, k, k, generator=g) * (2.0 / (Ci * k * k)) ** 0.5).contiguous()
        r = torch.randn(B, Co, HW, HW, generator=g).cuda() if res else None
        y = torch.empty(B, Co, HW, HW, device="cuda")
        for mname, env in MODES:
            if k == 3 and mname.startswith("astat"):
                continue
            cells = []
            for bn, epi in ((64, 1), (128, 1), (256, 1), (128, 2)):
                for kk in KEYS:
                    os.environ.pop(kk, None)
                os.environ.update(env)
                os.environ["YB_CONV2D_BN"] = str(bn)
                if epi == 2:
                    os.environ["YB_CONV2D_EPI"] = "2"
                ms = ctypes.c_float(0)
                st = lib.yb_conv2d(h, _lib.ptr(x), ctypes.c_void_p(w.data_ptr()), None, _lib.ptr(r), _lib.ptr(y), B, Ci, HW, HW, Co, k, k,
                                   1, k // 2, 1, prec, 30, ctypes.byref(ms), _lib.current_stream())
                cells.append("%.1f" % (ms.value * 1e3) if st == 0 else "-")
            print("| %s | %s | %s |" % (name, mname, " | ".join(cells)), flush=True)
