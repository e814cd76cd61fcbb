#!/usr/bin/env python
"""bench.py -- frames/sec of the YOLACT inference path (Yolact.forward -> Detect -> postprocess) on
synthetic 550x550 frames, yolact_base (ResNet101-FPN), batch 8 per GPU (BASELINE.json configs[1]).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm on the host CPU cores

One step = one batch through the whole path.  Prints ONE JSON line (rank 0).  See DESIGN.md
"Measurement" for what each key means; in short:
  value     frames/s, inputs resident in HBM, fp32 masks [n,550,550] written for every detection
  e2e       frames/s through the reference-facing API with pinned HOST inputs: H2D of the frames,
            forward + Detect + mask assembly, D2H of classes/scores/boxes/bit-packed masks, every step
  roofline  tcgen05 conv stack: algorithmic FLOPs / CUDA-event time of the
            conv-stack graph, against the measured sustained cuBLAS bf16 peak (MEASURED_PEAKS.json)
  cpu_baseline  the oracle port (torch-CPU fp32 conv stack + numpy Detect/postprocess) on a bounded
            sample of the same workload, all host threads
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic conv FLOPs per image, 2*MAC (counted from the reference model)
GFLOP_PER_IMAGE = {
    "yolact_resnet50_config": 118.28, "yolact_base_config": 164.68, "yolact_plus_resnet50_config": 141.38,
    "yolact_plus_base_config": 187.33, "yolact_im700_config": 262.93, "yolact_darknet53_config": 154.71,
}


# FPS the reference publishes for each config (Titan Xp, batch 1, fp32, `eval.py --benchmark`,
# README.md:70-80).  Other hardware and batch size, so it is reported as an informational ratio, not as `vs_baseline`.
PUBLISHED_TITAN_XP_FPS = {
    "yolact_resnet50_config": 42.5, "yolact_darknet53_config": 40.0, "yolact_base_config": 33.5,
    "yolact_im700_config": 23.6, "yolact_plus_resnet50_config": 33.5, "yolact_plus_base_config": 27.3,
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="yolact_base_config")
    ap.add_argument("--batch", type=int, default=8, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=0, help="image size (default: the config's max_size)")
    ap.add_argument("--precision", default="f16x3", choices=["f16x3", "f16tc", "f32"],
                    help="f16x3: split-precision tcgen05 (meets the reference tolerance; the headline mode); "
                         "f16tc: single-pass fp16 tcgen05 (fast mode); f32: CUDA-core fp32")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=10, help="images in the cpu_baseline sample")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
def host_threads():
    """Threads the CPU arm may use: affinity mask capped by the cgroup CPU quota (a container that sees
    128 CPUs but owns 16 of them must not spin 128 OpenMP threads)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, n)


def oracle_pipeline(cfg, sd):
    """The reference algorithm on the CPU (oracle port): net(x) + Detect + postprocess for one batch.
    torch-CPU fp32 conv stack (oracle/yolact_oracle.py) + torch-CPU Detect/postprocess
    (oracle/torch_port.py): the same ATen kernels the reference graph runs with --cuda=False."""
    import torch
    from oracle import yolact_oracle as O
    from oracle import torch_port as T
    orc = O.ConvStackOracle(cfg, sd)

    def run(x, out_hw):
        with torch.no_grad():
            raw = orc.forward(x)
            conf = torch.softmax(raw["conf"], -1)
            n_det = 0
            for b in range(x.shape[0]):
                det = T.detect_one(raw["loc"][b], conf[b], raw["mask"][b], raw["priors"], cfg.nms_conf_thresh,
                                   cfg.nms_thresh, cfg.nms_top_k, cfg.max_num_detections)
                if det is None:
                    continue
                det["proto"] = raw["proto"][b]
                fn = orc.maskiou if cfg.use_maskiou else None
                classes, scores, boxes, masks = T.postprocess_one(det, out_hw[1], out_hw[0], maskiou_fn=fn)
                n_det += int(masks.shape[0])
        return n_det
    return run


def pick_threads(run, size):
    """Best-performing thread count for the CPU arm (more threads is not always faster on a big host)."""
    import torch
    from oracle.weights import deterministic_input
    limit = host_threads()
    cands = sorted(set(min(c, limit) for c in (8, 16, 32, 64, 128, limit)))
    x = deterministic_input(1, size, size, 31337)
    best, best_t = cands[0], 1e30
    for c in cands:
        torch.set_num_threads(c)
        run(x, (size, size))
        t0 = time.perf_counter()
        run(x, (size, size))
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
        elif dt > 1.5 * best_t:
            break   # past the knee: more threads only add synchronisation cost
    torch.set_num_threads(best)
    return best


def time_cpu(cfg, sd, size, n_images, seed=4321):
    import torch
    from oracle.weights import deterministic_input
    run = oracle_pipeline(cfg, sd)
    pick_threads(run, size)
    x = deterministic_input(1, size, size, seed)
    run(x, (size, size))  # warm-up (thread pools, allocator)
    t0 = time.perf_counter()
    nd = 0
    for i in range(n_images):
        nd += run(deterministic_input(1, size, size, seed + 1 + i), (size, size))
    dt = time.perf_counter() - t0
    return n_images / dt, nd / max(1, n_images)


# ---------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    from yolact_b200.config import CONFIGS
    from oracle.weights import deterministic_state_dict, deterministic_input

    cfg = CONFIGS[args.config].copy()
    size = args.size or cfg.max_size
    B = args.batch
    workload = "%s @%d, batch %d/GPU, synthetic frames, random-init deterministic weights (100 detections/image)" % (
        cfg.name, size, B)
    base = {
        "metric": "frames/sec @ %dx%d %s (Yolact.forward + Detect + postprocess)" % (size, size, cfg.name),
        "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic",
    }

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        import yolact_b200
        net = yolact_b200.Yolact(cfg)  # parameter holder only: gives the reference's state_dict keys
        sd = deterministic_state_dict(net.state_dict(), 0)
        run = oracle_pipeline(cfg, sd)
        pick_threads(run, size)
        per_step = 1   # bounded sample: 1 image of the same workload per step
        for i in range(max(1, min(args.warmup, 2))):
            run(deterministic_input(per_step, size, size, 7000 + i), (size, size))
        t0 = time.perf_counter()
        for i in range(args.steps):
            run(deterministic_input(per_step, size, size, 8000 + i), (size, size))
        dt = time.perf_counter() - t0
        fps = per_step * args.steps / dt
        line = dict(base)
        line.update({
            "impl": "reference", "value": fps, "ms_per_step": 1e3 * dt / args.steps, "dtype": "f32", "n_gpus": args.gpus,
            "config": {"workload": workload, "sample": "%d image(s) of the workload per step" % per_step},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "%d steps x %d image(s), oracle port (torch-CPU fp32 conv stack + torch-CPU "
                                       "Detect/postprocess); /root/reference does not exist on the GPU box; threads "
                                       "chosen by timing (host limit %d)" % (args.steps, per_step, host_threads())},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        })
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch.distributed as dist
    import yolact_b200
    from yolact_b200.output_utils import assemble_masks_batch
    from yolact_b200 import output_utils
    from yolact_b200.parallel import gather_detections

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    yolact_b200.cfg.replace(cfg.copy())
    net = yolact_b200.Yolact(cfg, precision=args.precision)
    sd = deterministic_state_dict(net.state_dict(), 0)
    net.load_state_dict(sd)
    net.eval()

    n_rot = 6  # 6 distinct input batches (6 x 29 MB = 174 MB > 126 MB L2) rotate through the loop
    xs = [deterministic_input(B, size, size, 1234 + 100 * rank + i).to(dev) for i in range(n_rot)]
    M, k = cfg.max_num_detections, cfg.mask_dim
    # Two compute streams: the mask assembly of step i (HBM-write bound) runs on s_post while the conv stack of step
    # i+1 (tensor / L2 bound) already runs on the main stream; outputs are double-buffered.  BENCH_OVERLAP=0 serialises.
    overlap = os.environ.get("BENCH_OVERLAP", "1") != "0"
    masks_f32 = [torch.empty(B, M, size, size, dtype=torch.float32, device=dev) for _ in range(2 if overlap else 1)]
    s_post = torch.cuda.Stream() if overlap else None
    # The network outputs of step i are read by s_post; they are kept referenced until step i+2, whose first action
    # is to make the main stream wait for step i's mask assembly -- so the caching allocator (which only tracks the
    # allocating stream) can never hand their memory to main-stream work that runs before s_post has read it.
    hold = [None, None]
    ev_post_v = [torch.cuda.Event(), torch.cuda.Event()]

    def step_device(i, fmt="f32", out=None):
        if overlap and hold[i % 2] is not None:
            torch.cuda.current_stream().wait_event(ev_post_v[i % 2])
        box, coef, cls, score, count, proto = net.infer_padded(xs[i % n_rot])
        # all M padded rows are assembled (no host sync on the count); with these weights count == M
        if overlap:
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(s_post):
                s_post.wait_event(ev)
                res = assemble_masks_batch(proto, coef, box, size, size, True, fmt, masks_out=out[i % 2])
                ev_post_v[i % 2].record(s_post)
            hold[i % 2] = (box, coef, proto, res)
        else:
            res = assemble_masks_batch(proto, coef, box, size, size, True, fmt, masks_out=out[0])
        if world > 1:
            gather_detections(box, coef, cls, score, count, per_rank_batch=B)
        return box, coef, cls, score, count, res

    def barrier_sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier_sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        if s_post is not None:
            torch.cuda.current_stream().wait_stream(s_post)   # the last steps' mask assembly is inside the timed region
        e1.record()
        barrier_sync()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- value: inputs resident in HBM, fp32 masks
    l0 = net.launch_count() + output_utils.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(lambda i: step_device(i, "f32", masks_f32), args.steps, max(3, args.warmup))
    clocks = sampler.stop() if rank == 0 else None
    launches = (net.launch_count() + output_utils.launch_count() - l0)
    fps = world * B * args.steps / (ms / 1e3)

    # ---- e2e: pinned host inputs -> H2D -> path -> D2H of classes/scores/boxes/bit-packed masks.
    # Three streams, double-buffered: the H2D of step i+1 and the D2H of step i-1 overlap the compute of
    # step i.  Every step's inputs come from pinned host memory and every step's result is read on the
    # host (the loop blocks on step i-1's D2H event before issuing step i+1).
    wpr = (size + 31) // 32
    hx = [deterministic_input(B, size, size, 555 + i).pin_memory() for i in range(2)]
    dx = [torch.empty(B, 3, size, size, device=dev) for _ in range(2)]
    h_cls = [torch.empty(B, M, dtype=torch.int64).pin_memory() for _ in range(2)]
    h_score = [torch.empty(B, M, dtype=torch.float32).pin_memory() for _ in range(2)]
    h_count = [torch.empty(B, dtype=torch.int32).pin_memory() for _ in range(2)]
    h_boxes = [torch.empty(B, M, 4, dtype=torch.int64).pin_memory() for _ in range(2)]
    h_masks = [torch.empty(B, M, size, wpr, dtype=torch.int32).pin_memory() for _ in range(2)]
    d_masks = [torch.empty(B, M, size, wpr, dtype=torch.int32, device=dev) for _ in range(2)]
    d_boxes = [torch.empty(B, M, 4, dtype=torch.int64, device=dev) for _ in range(2)]
    h2d = B * 3 * size * size * 4
    d2h = sum(t[0].numel() * t[0].element_size() for t in (h_cls, h_score, h_count, h_boxes, h_masks))
    s_main = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_comp = [torch.cuda.Event() for _ in range(2)]
    ev_post = [torch.cuda.Event() for _ in range(2)]
    hold_e = [None, None]
    ev_d2h = [torch.cuda.Event() for _ in range(2)]
    state = {"n": 0}

    def step_e2e(_i):
        i = state["n"]
        state["n"] += 1
        k = i % 2
        with torch.cuda.stream(s_in):
            if i >= 2:
                s_in.wait_event(ev_comp[k])          # dx[k] was last read by the compute of step i-2
            dx[k].copy_(hx[k], non_blocking=True)
            ev_in[k].record(s_in)
        s_main.wait_event(ev_in[k])
        if i >= 2 and overlap:
            s_main.wait_event(ev_post[k])            # step i-2's mask assembly has read its inputs (see `hold`)
        if i >= 2:
            # d_masks[k] / d_boxes[k] were last read by step i-2's D2H
            (s_post if overlap else s_main).wait_event(ev_d2h[k])
        box, coef, cls, score, count, proto = net.infer_padded(dx[k])
        ev_comp[k].record(s_main)                    # dx[k] is free again once the network has read it
        if overlap:
            with torch.cuda.stream(s_post):
                s_post.wait_event(ev_comp[k])
                assemble_masks_batch(proto, coef, box, size, size, True, "bits", masks_out=d_masks[k], boxes_out=d_boxes[k])
                ev_post[k].record(s_post)
            hold_e[k] = (box, coef, proto)
        else:
            assemble_masks_batch(proto, coef, box, size, size, True, "bits", masks_out=d_masks[k], boxes_out=d_boxes[k])
            ev_post[k].record(s_main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_post[k])
            for t in (cls, score, count):
                t.record_stream(s_out)
            h_cls[k].copy_(cls, non_blocking=True)
            h_score[k].copy_(score, non_blocking=True)
            h_count[k].copy_(count, non_blocking=True)
            h_boxes[k].copy_(d_boxes[k], non_blocking=True)
            h_masks[k].copy_(d_masks[k], non_blocking=True)
            ev_d2h[k].record(s_out)
        if i >= 1:
            ev_d2h[(i - 1) % 2].synchronize()        # step i-1's result is on the host

    def e2e_drain():
        ev_d2h[(state["n"] - 1) % 2].synchronize()

    for i in range(3):
        step_e2e(i)
    e2e_drain()
    barrier_sync()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_e2e(i)
    e2e_drain()
    s_main.wait_stream(s_out)
    if s_post is not None:
        s_main.wait_stream(s_post)
    e1.record()
    barrier_sync()
    ms_e2e = e0.elapsed_time(e1)
    wall_e2e = (time.perf_counter() - t0) * 1e3
    if world > 1:
        tt = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_e2e = float(tt.item())
    fps_e2e = world * B * args.steps / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel family (tcgen05 conv stack), timed live with CUDA events
    net.train()
    def fwd_only(i):
        net.forward_conv_only(xs[i % n_rot])
    ms_conv = timed(fwd_only, args.steps, 3)
    net.eval()
    gflop = GFLOP_PER_IMAGE.get(args.config, 0.0) * size * size / float(cfg.max_size * cfg.max_size)
    achieved = gflop * B * args.steps / (ms_conv / 1e3) / 1e3   # TFLOP/s per GPU
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "measured sustained cuBLAS bf16 (MEASURED_PEAKS.json)" if peaks else "fallback ~1.4 PF sustained (B200_PROFILING.md)"
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("traffic_bytes_per_step")
    except Exception:
        pass
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "tc_conv_kernel<BN> (all conv launches of one step)",
                "ms_conv_stack_per_step": ms_conv / args.steps, "peak_source": peak_src,
                "algorithmic_gflop_per_step": gflop * B}

    line = dict(base)
    line.update({
        "value": fps, "ms_per_step": ms / args.steps, "dtype": {"f16x3": "f16x3 (fp16 hi+lo operand pairs, 3 tcgen05 passes, fp32 accumulate: fp32-equivalent)",
                                                          "f16tc": "f16", "f32": "f32"}[args.precision],
        "config": {"workload": workload, "global_batch": world * B, "image_size": size, "parallelism": "dp%d" % world,
                   "detections_per_image": M, "mask_format_value": "f32 [n,h,w]", "mask_format_e2e": "1 bit/pixel",
                   "l2": "6 rotating input batches (174 MB) and ~2 GB of activations+masks per step exceed the 126 MB L2",
                   "cuda_graph": True,
                   "streams": "mask assembly of step i overlaps the conv stack of step i+1 (2 compute streams)" if overlap
                              else "single compute stream"},
        "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps, "host_wall_ms_per_step": wall_e2e / args.steps,
                "pipelining": "double-buffered streams: H2D(i+1) | network(i+1) | mask assembly(i) | D2H(i-1); host blocks on each step's D2H"},
        "gpu_launches": int(launches),
        "vs_published_titan_xp": {"ratio_per_gpu": fps / world / PUBLISHED_TITAN_XP_FPS[args.config],
                                  "published_fps": PUBLISHED_TITAN_XP_FPS[args.config],
                                  "note": "reference README (BASELINE.md section 1): 1 Titan Xp, batch 1, fp32"}
        if args.config in PUBLISHED_TITAN_XP_FPS and size == cfg.max_size else None,
        "clocks": clocks,
        "roofline": roofline,
    })

    # ---- cpu_baseline (rank 0, N == 1 only): bounded sample of the same workload on the host cores
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_fps, nd = time_cpu(cfg, sd, size, args.cpu_sample)
        line["cpu_baseline"] = {"value": cpu_fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "%d images of the same workload (batch 1), %.0f detections/image" % (args.cpu_sample, nd)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
