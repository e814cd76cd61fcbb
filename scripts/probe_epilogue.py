"""Single-layer timings of the epilogue-buffering variants (one / two epilogue groups, N tile) through yb_conv2d.
HISTORICAL: the three-buffer variant (YB_CONV2D_EPI=3) it was written for lost (profiles/r2_call10_summary.txt) and was
removed; on the current library EPI=3 is treated as one group."""
import ctypes, os, sys
sys.path.insert(0, "/root/repo")
import torch
from yolact_b200 import _lib
lib = _lib.load()
yc = _lib.YbConfig(); yc.backbone = _lib.YB_BACKBONE_NONE
yc.num_classes, yc.mask_dim, yc.precision = 81, 32, _lib.YB_PREC_F32
yc.nms_top_k, yc.nms_conf_thresh, yc.nms_thresh, yc.max_num_detections = 200, 0.05, 0.5, 100
h = ctypes.c_void_p(); _lib.check(lib.yb_create(ctypes.byref(yc), 0, ctypes.byref(h)), "create")
g = torch.Generator().manual_seed(0)
def run(prec, B, Ci, HW, Co, k, res, env):
    for kk in ["YB_CONV2D_PAIR", "YB_CONV2D_SK", "YB_CONV2D_BN", "YB_CONV2D_EPI", "YB_CONV2D_GRID"]:
        os.environ.pop(kk, None)
    os.environ.update(env)
    x = torch.randn(B, Ci, HW, HW, generator=g).cuda()
    w = (torch.randn(Co, Ci, k, k, generator=g) * (2.0 / (Ci * k * k)) ** 0.5).contiguous()
    r = torch.randn(B, Co, HW, HW, generator=g).cuda() if res else None
    y = torch.empty(B, Co, HW, HW, device="cuda")
    ms = ctypes.c_float(0)
    st = lib.yb_conv2d(h, _lib.ptr(x), ctypes.c_void_p(w.data_ptr()), None, _lib.ptr(r), _lib.ptr(y), B, Ci, HW, HW, Co, k, k, 1, k // 2, 1, prec, 30, ctypes.byref(ms), _lib.current_stream())
    return ms.value * 1e3 if st == 0 else -1
for prec in (3, 1):
    print("precision", prec)
    for res in (True, False):
        for Ci, Co in ((256, 1024), (64, 1024), (256, 256), (1024, 1024)):
            row = []
            for env in ({"YB_CONV2D_BN": "64"}, {"YB_CONV2D_BN": "128"}, {"YB_CONV2D_BN": "256"}, {"YB_CONV2D_BN": "128", "YB_CONV2D_EPI": "2"}, {"YB_CONV2D_BN": "256", "YB_CONV2D_EPI": "2"},
                        {"YB_CONV2D_BN": "64", "YB_CONV2D_EPI": "3"}, {"YB_CONV2D_BN": "128", "YB_CONV2D_EPI": "3"}):
                row.append("%.1f" % run(prec, 8, Ci, 35, Co, 1, res, env))
            print("  1x1 %4d->%4d @35^2 res=%d | bn64 %s bn128 %s bn256 %s bn128e2 %s bn256e2 %s bn64e3 %s bn128e3 %s" % ((Ci, Co, res) + tuple(row)), flush=True)
