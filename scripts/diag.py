import faulthandler, os, sys, time
faulthandler.dump_traceback_later(45, repeat=True, file=sys.stderr)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t0 = time.time()
import torch, yolact_b200
print("imported %.1fs" % (time.time() - t0), flush=True)
from oracle.weights import deterministic_state_dict, deterministic_input
from yolact_b200.config import CONFIGS
cfgname = sys.argv[1] if len(sys.argv) > 1 else "yolact_resnet50_config"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 160
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cfg = CONFIGS[cfgname].copy(); yolact_b200.cfg.replace(cfg.copy())
net = yolact_b200.Yolact(cfg); net.load_state_dict(deterministic_state_dict(net.state_dict(), 0)); net.eval(); net.detect.use_fast_nms = True
print("net built %.1fs" % (time.time() - t0), flush=True)
x = deterministic_input(B, size, size, 1).cuda()
torch.cuda.synchronize()
print("input on gpu %.1fs" % (time.time() - t0), flush=True)
h = net._handle_for(x.device)
print("weights pushed + finalized %.1fs" % (time.time() - t0), flush=True)
net.train()
o = net.forward_conv_only(x); torch.cuda.synchronize()
print("conv stack eager ok %.1fs" % (time.time() - t0), flush=True)
net.eval()
for i in range(4):
    t = time.time(); out = net.infer_padded(x); torch.cuda.synchronize()
    print("call", i, "%.3f s" % (time.time() - t), "count", out[4].tolist(), flush=True)
