#!/usr/bin/env python
"""bench.py -- scenes/s, forward+backward, of the scene-graph hot path on N B200s (BASELINE.json metric).

  python bench.py [--gpus N --steps K --warmup W]            sg4d (this repo) on the GPU(s)
  python bench.py --impl reference [--gpus N ...]            the reference arithmetic on the host CPU

One "step" = one forward + loss + backward of the full pipeline (both PointNet++ encoders, TripletGCN,
heads, weighted NLL loss; gradient all-reduce when N > 1) over one batch of synthetic scenes
(BASELINE configs[2]: 8 scenes x (12 object + 66 edge clouds) x 80 000 points per GPU).  Rank 0 prints
ONE JSON line; see DESIGN.md "Measurement" for every field.

The reference has no CPU implementation of its custom ops (EXT/src/*.cpp: "CPU not supported"), so the
reference arm and `cpu_baseline` time the oracle port (oracle/pn2_oracle.c + oracle/model_ref.py: the
reference's op sequence in stock PyTorch CPU ops) on the box's host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CLOUDS_PER_SCENE = 78  # 12 objects + 66 edges

# stdout carries exactly ONE JSON line: everything else that native libraries print there (NCCL's version banner, ...)
# is routed to stderr by pointing fd 1 at fd 2 for the duration of the run
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sg4d", choices=["sg4d", "reference"])
    ap.add_argument("--scenes-per-gpu", type=int, default=8)
    ap.add_argument("--points", type=int, default=80000, help="points per cloud (BL = 80000)")
    ap.add_argument("--points-rel", type=int, default=None, help="points per edge cloud (default: --points)")
    ap.add_argument("--n-obj", type=int, default=12)
    ap.add_argument("--pairs", default="unordered", choices=["unordered", "ordered"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="run the two encoders on one stream")
    ap.add_argument("--cpu-clouds", type=int, default=78, help="clouds in the CPU legs (default: one full scene; fewer = a scaled sample)")
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json configs[config-1]: 1 = 1 scene x 2048 pts forward only; 2 = 8 scenes, encoders only, "
                         "fwd+bwd; 3 = 8 scenes, full pipeline (default, the headline); 4 = 32 scenes + image branch; "
                         "5 = 32 scenes per GPU (256 at 8 GPUs)")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"], help="tensor-core operand precision (config 4: bf16)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying its CUDA graph")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip timing the reference's own code on the GPU")
    ap.add_argument("--e2e-input", default="scene", choices=["scene", "crops", "both"],
                    help="what the end-to-end leg uploads every step: raw scenes (GPU front-end builds the crops) or the crops")
    ap.add_argument("--scene-points", type=int, default=200000, help="points per raw scene of the end-to-end leg")
    a = ap.parse_args()
    a.forward_only, a.encoders_only, a.image = False, False, False
    a.bf16 = a.config == 4 or a.precision == "bf16"
    if a.config == 1:
        a.scenes_per_gpu, a.points, a.n_obj, a.forward_only = 1, 2048, 4, True
    elif a.config == 2:
        a.encoders_only = True
    elif a.config == 4:
        a.scenes_per_gpu, a.image = 32, True
    elif a.config == 5:
        a.scenes_per_gpu = 32
    return a


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def model_config():
    return json.load(open(os.path.join(ROOT, "4d-or_b200", "default_config.json")))


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU arm

def cpu_sample_batch(args, n_clouds):
    """The workload of the CPU legs: ONE full scene of the benchmark shape (12 + 66 clouds, its real edge list) when
    n_clouds covers it; otherwise a bounded sample -- n_clouds clouds in the scene's own object:edge proportion, the
    GNN run on as many (random) edges as there are edge clouds -- whose scenes/s is scaled by the cloud count."""
    from sg4d import synthetic
    pr = args.points_rel or args.points
    n_edge = args.n_obj * (args.n_obj - 1) // (2 if args.pairs == "unordered" else 1)
    if n_clouds >= args.n_obj + n_edge:
        return synthetic.make_scene(4321, n_obj=args.n_obj, n_points_obj=args.points, n_points_rel=pr, pairs=args.pairs), \
            args.n_obj + n_edge, "one full scene"
    n_obj = max(2, round(n_clouds * args.n_obj / (args.n_obj + n_edge)))   # BatchNorm1d over the nodes / edges needs >= 2 rows
    n_rel = max(2, n_clouds - n_obj)
    gen = torch.Generator().manual_seed(4321)
    obj = torch.stack([synthetic.make_cloud(gen, args.points, 6) for _ in range(n_obj)])
    rel = torch.stack([synthetic.make_cloud(gen, pr, 7) for _ in range(n_rel)])
    ei = torch.stack([torch.randint(0, n_obj, (n_rel,), generator=gen), torch.randint(0, n_obj, (n_rel,), generator=gen)])
    one_hot = torch.zeros(n_rel, 12)
    one_hot[:, 0] = 1
    one_hot[:, 6] = 1
    return {"obj_points": obj.permute(0, 2, 1), "rel_points": rel.permute(0, 2, 1), "edge_indices": ei,
            "relation_objects_one_hot": one_hot, "gt_class": torch.randint(0, 12, (n_obj,), generator=gen),
            "gt_rels": torch.randint(0, 15, (n_rel,), generator=gen)}, n_obj + n_rel, \
        f"SAMPLE of {n_obj + n_rel} of a scene's {args.n_obj + n_edge} clouds with random edges, scaled by cloud count"


def cpu_step(sd, batch, args):
    from oracle import model_ref
    for v in sd.values():
        if v.requires_grad:
            v.grad = None
    if args.forward_only:
        with torch.no_grad():
            outs = model_ref.forward(sd, batch, training=True, dropout=True)
            return float(model_ref.loss_fn(outs[0], outs[1], batch, torch.ones(12), torch.ones(15), 1e-6))
    if args.encoders_only:
        loss = model_ref.encoder(sd, "obj_encoder", batch["obj_points"], True).sum() + \
            model_ref.encoder(sd, "rel_encoder", batch["rel_points"], True).sum()
    else:
        outs = model_ref.forward(sd, batch, training=True, dropout=True)
        loss = model_ref.loss_fn(outs[0], outs[1], batch, torch.ones(12), torch.ones(15), 1e-6)
    loss.backward()
    return float(loss.detach())


def run_cpu(args, steps, warmup, budget_s):
    """Times the oracle port on the host cores; returns (scenes/s, ms per step, description dict)."""
    from oracle import model_ref, pn2_ext_cpu, weights
    pn2_ext_cpu.build()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    sd = model_ref.clone_state(weights.synth_state_dict(seed=0))
    n_edge = args.n_obj * (args.n_obj - 1) // (2 if args.pairs == "unordered" else 1)
    scene_clouds = args.n_obj + n_edge
    n_clouds = min(args.cpu_clouds, scene_clouds)
    while True:
        batch, used, what = cpu_sample_batch(args, n_clouds)
        t0 = time.perf_counter()
        cpu_step(sd, batch, args)
        t1 = time.perf_counter() - t0
        if t1 * (steps + warmup) <= budget_s or n_clouds <= 2:
            break
        n_clouds = max(2, int(n_clouds * budget_s / (t1 * (steps + warmup))))
    for _ in range(max(0, warmup - 1)):
        cpu_step(sd, batch, args)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_step(sd, batch, args)
        times.append(time.perf_counter() - t0)
    per_step = sum(times) / len(times)
    scenes = used / scene_clouds
    desc = {"kind": "port", "cores": cores, "unit": "scenes/s",
            "sample": f"{what} ({args.points} pts per cloud), {'forward' if args.forward_only else 'fwd+loss+bwd'}, "
                      f"{steps} timed steps of {per_step:.2f} s; oracle/pn2_oracle.c (OpenMP over clouds) + "
                      f"oracle/model_ref.py (torch CPU, {cores} threads)"}
    return scenes / per_step, per_step * 1e3, desc


def workload_name(args):
    pr = args.points_rel or args.points
    n_edge = args.n_obj * (args.n_obj - 1) // (2 if args.pairs == "unordered" else 1)
    what = {1: "full pipeline, forward only", 2: "PointNet++ MSG encoders only (upstream gradient = ones), fwd+bwd",
            3: "PointNet++ MSG encoders + TripletGCN + heads, fwd+loss+bwd",
            4: "full pipeline + image-feature concat branch, fwd+loss+bwd, bf16 operands / fp32 accumulation",
            5: "full pipeline, fwd+loss+bwd, scene-sharded data parallel"}[args.config]
    return (f"BASELINE configs[{args.config - 1}]: {args.scenes_per_gpu} scenes/GPU x ({args.n_obj} obj x {args.points} pts + "
            f"{n_edge} edges x {pr} pts), {what}, {'bf16' if args.bf16 else 'fp32'}")


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, ms, desc = run_cpu(args, args.steps, args.warmup, budget_s=200.0)
    desc["value"] = value
    line = {"impl": "reference", "metric": "scenes/sec fwd+bwd", "value": value, "unit": "scenes/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "cpu_sample": desc["sample"],
                       "note": "reference has no CPU path for its custom ops; this is the oracle port of its arithmetic on "
                               "the host cores, rank 0 only; scenes/s of the CPU sample (see cpu_sample)"},
            "cpu_baseline": desc,
            "e2e": {"value": value, "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------------------------- GPU arm

def main_sg4d(args):
    import torch.distributed as dist
    import sg4d
    from sg4d import _lib, parallel, synthetic
    from sg4d.model import SGPNModelWrapper

    rank, world, local_rank = parallel.init_from_env()
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl sg4d needs a CUDA device (there is no CPU fallback)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    assert _lib.load().sg4d_check_device() == 0, "sg4d needs an sm_100 (B200) device"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    if args.bf16:
        _lib.set_precision("bf16")
    cfg = model_config()
    if args.image:
        cfg["IMAGE_INPUT"] = "full"
    torch.manual_seed(0)
    names = [f"rel{i}" for i in range(14)] + ["none"]
    model = SGPNModelWrapper(cfg, 12, 15, torch.ones(12), torch.ones(15), names).to(dev).train()
    model.overlap_encoders = not args.no_overlap
    bucket = parallel.GradBucket(model)

    S = args.scenes_per_gpu
    pr = args.points_rel or args.points
    host = synthetic.make_batch(rank * S, S, n_obj=args.n_obj, n_points_obj=args.points, n_points_rel=pr,
                                pairs=args.pairs, image=args.image)
    tensor_keys = [k for k, v in host.items() if torch.is_tensor(v)]
    pinned = {}
    for k in tensor_keys:
        v = host[k]
        if k.endswith("_points"):
            v = v.permute(0, 2, 1).contiguous().pin_memory().permute(0, 2, 1)
        else:
            v = v.contiguous().pin_memory()
        pinned[k] = v
    h2d_bytes = sum(v.numel() * v.element_size() for v in pinned.values())
    resident = synthetic.to_device(pinned, dev)

    def step(batch):
        if args.forward_only:                       # config 1: forward + loss, no gradients
            with torch.no_grad():
                return model.training_step(batch)
        bucket.zero()
        if args.encoders_only:                      # config 2: both encoders, upstream gradient of ones, no GNN / heads
            obj_f, rel_f = model._encode(batch)
            loss = obj_f.sum() + rel_f.sum()
        else:
            loss = model.training_step(batch)
        loss.backward()
        bucket.all_reduce_mean()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eager_step = step
    graphed = None
    if not args.no_graph:
        # the whole forward + loss + backward is captured once and replayed: one graph launch per step instead of ~800 launches
        from sg4d import graph as sg_graph

        def body(b):
            if args.forward_only:
                with torch.no_grad():
                    return model.training_step(b)
            if args.encoders_only:
                obj_f, rel_f = model._encode(b)
                return obj_f.sum() + rel_f.sum()
            return model.training_step(b)

        graphed = sg_graph.GraphedStep(model, resident, bucket, step_fn=body)

        def step(batch):
            loss = graphed(batch)
            if not args.forward_only:
                bucket.all_reduce_mean()
            return loss

    for _ in range(args.warmup):
        step(resident)
    barrier()

    # ---- timed region 1: inputs resident in HBM (working set of 1.4 GB/step >> 126 MB L2)
    launches0 = _lib.LAUNCH_COUNT
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with ClockSampler(local_rank) as clk:
        barrier()
        ev[0].record()
        for i in range(args.steps):
            step(resident)
            ev[i + 1].record()
        barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    launches = _lib.LAUNCH_COUNT - launches0
    peak_mem_gb = torch.cuda.max_memory_reserved(dev) / 1e9      # reserved: the captured graph's private pool counts

    # ---- kernel table: the same steps once more with a CUDA-event pair around every C-ABI call (on the launching
    #      stream).  The encoders share one stream here so that a call's duration is its own, not that of whatever
    #      ran beside it.
    model.overlap_encoders = False
    _lib.enable_timing(True)
    _lib.drain_timing()
    barrier()
    kt0, kt1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lc0 = _lib.LAUNCH_COUNT
    kt0.record()
    for i in range(args.steps):
        eager_step(resident)
    kt1.record()
    if graphed is not None:      # a replayed graph launches the same kernels as the eager step it was captured from
        launches = _lib.LAUNCH_COUNT - lc0
    barrier()
    per_call = _lib.drain_timing()
    _lib.enable_timing(False)
    table_ms_per_step = kt0.elapsed_time(kt1) / args.steps
    model.overlap_encoders = not args.no_overlap
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * S / (ms_per_step / 1e3)

    # ---- timed region 2: end to end through the public API, host (pinned) buffers in, loss out.
    #      Default input = RAW SCENES (points + per-point object masks, --scene-points each): every step uploads its S scenes
    #      (a few MB each) from pinned host memory and the GPU front-end (sg4d.frontend: crops, union boxes, sampling, zero_mean)
    #      builds the 12 + 66 clouds of --points points per scene on the device -- the step immediately before the hot path
    #      in the reference (SGH/dataset/data_preparation_utils.py), there on the CPU.  --e2e-input crops uploads the
    #      pre-cropped clouds instead (171 MB per scene; round 1's path).
    def timed_e2e(make_batch, h2d, after_launch=None):
        """make_batch(i): step i's inputs (device tensors whose upload / preparation was queued earlier on a side stream);
        after_launch(i): queues step i + 1's upload / preparation -- called AFTER step i has been launched, so that the host time
        of those launches and their execution overlap step i on the GPU."""
        for i in range(2):                      # untimed: allocations
            step(make_batch(i))
            if after_launch:
                after_launch(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            loss = step(make_batch(i))
            if after_launch:
                after_launch(i)
            loss.detach().to("cpu", non_blocking=False)       # loss read back every step
        e1.record()
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return {"value": world * S / (float(tt.item()) / args.steps / 1e3), "unit": "scenes/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4}

    e2e = e2e_crops = None
    if not args.no_e2e:
        if args.e2e_input in ("scene", "both"):
            from sg4d import frontend
            raw = [synthetic.make_raw_scene(rank * S + i, n_obj=args.n_obj, n_points=args.scene_points) for i in range(S)]
            raw_pts = torch.stack([r[0] for r in raw]).pin_memory()           # (S, P, 6)
            raw_msk = torch.stack([r[1] for r in raw]).pin_memory()           # (S, P)
            small = {k: pinned[k] for k in ("relation_objects_one_hot", "gt_class", "gt_rels") + (("full_image_features",) if args.image else ())}
            side = torch.cuda.Stream(device=dev)
            bufs = [None, None]

            E_ = resident["rel_points"].shape[0] // S

            def stage(slot):
                """one step's scenes: upload AND crop on the side stream while the previous step computes on the main stream (the
                front-end's kernels are small latency-bound grids without shared memory: they co-reside with the MLP kernels).
                Every scene's crops are written straight into its slices of the batch tensors."""
                with torch.cuda.stream(side):
                    pts_d, msk_d = raw_pts.to(dev, non_blocking=True), raw_msk.to(dev, non_blocking=True)
                    sm = {k: v.to(dev, non_blocking=True) for k, v in small.items()}
                    obj_all = torch.empty(S * args.n_obj, args.points, raw_pts.shape[2], dtype=torch.float32, device=dev)
                    rel_all = torch.empty(S * E_, pr, raw_pts.shape[2] + 1, dtype=torch.float32, device=dev)
                    for k in range(S):
                        frontend.prepare_scene(pts_d[k], msk_d[k], args.n_obj, args.points, pr, pairs=args.pairs,
                                               out_obj=obj_all[k * args.n_obj:(k + 1) * args.n_obj], out_rel=rel_all[k * E_:(k + 1) * E_])
                    b = {"obj_points": obj_all.permute(0, 2, 1), "rel_points": rel_all.permute(0, 2, 1),
                         "edge_indices": resident["edge_indices"]}
                    b.update(sm)
                    if "edge_scene" in resident:
                        b["edge_scene"] = resident["edge_scene"]
                    ev = torch.cuda.Event()
                    ev.record(side)
                bufs[slot] = (b, ev)

            def scene_batch(i):
                if bufs[i & 1] is None:
                    stage(i & 1)
                b, ev = bufs[i & 1]
                main = torch.cuda.current_stream(dev)
                main.wait_event(ev)
                for v in b.values():
                    if torch.is_tensor(v):
                        v.record_stream(main)
                return b

            def stage_next(i):
                # slot (i + 1) & 1 was last read by step i - 1, whose loss the host has already read back: nothing to wait for
                stage((i + 1) & 1)

            h2d_scene = raw_pts.numel() * 4 + raw_msk.numel() * 4 + sum(v.numel() * v.element_size() for v in small.values())
            e2e = timed_e2e(scene_batch, h2d_scene, stage_next)
            e2e["input"] = f"{S} raw scenes of {args.scene_points} points + object masks per step; crops built by the GPU front-end"
            bufs = [None, None]
        if args.e2e_input in ("crops", "both"):
            pf = parallel.DevicePrefetcher(dev)
            state = {"primed": False}

            def crops_batch(i):
                if not state["primed"]:
                    pf.submit(pinned)
                    state["primed"] = True
                b = pf.next()
                pf.submit(pinned)
                return b

            r = timed_e2e(crops_batch, h2d_bytes)
            r["input"] = "pre-cropped clouds uploaded every step"
            if e2e is None:
                e2e = r
            else:
                e2e_crops = r

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel table and the roofline of the dominant own kernel
    peak, peak_src = load_peaks()
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tpeak = pk.get("bf16_tflops_sustained", 1384.6)
    kernels = []
    for (name, key), ms in per_call.items():
        kernels.append({"call": name, "args": list(key), "launches_per_step": len(ms) / args.steps,
                        "avg_ms": sum(ms) / len(ms), "ms_per_step": sum(ms) / args.steps})
    kernels.sort(key=lambda k: -k["ms_per_step"])
    own_ms = sum(k["ms_per_step"] for k in kernels)

    # Algorithmic bytes (inputs read once + outputs written once; a gathered cloud counts once per launch) and
    # algorithmic FLOPs (2 x MACs of the layer the call evaluates; recomputation is NOT counted) per call.
    # args = the call's small integer arguments in order (see include/sg4d.h).
    def model_of(nm, a):
        if nm in ("sg4d_fps_rows", "sg4d_fps_indexed"):
            return a[0] * (12 * a[1] + 16 * a[2]), 0
        if nm == "sg4d_spatial_index_build":
            return a[0] * a[1] * (12 + 20), 0
        if nm == "sg4d_group_rows":
            b_, n_, m_, ns_, c_ = a[:5]
            return b_ * (4 * m_ * ns_ + 4 * (3 + c_) * m_ * ns_ + 12 * m_) + b_ * min(n_, m_ * ns_) * 4 * (3 + c_), 0
        if nm == "sg4d_group_rows_grad":
            b_, n_, m_, ns_, c_ = a[:5]
            return b_ * (4 * c_ * m_ * ns_ + 4 * m_ * ns_ + 4 * c_ * n_), 0
        if nm == "sg4d_group_rows_grad_dy":      # (b, n, m, ns, c1): read idx, y1, dz1 once; write the per-point sums
            b_, n_, m_, ns_, c_ = a[:5]
            return b_ * (m_ * ns_ * (4 + 8 * c_) + 4 * c_ * n_), 0
        if nm == "sg4d_gather_y1":               # (rows, n, m, ns, c1): read idx, Z (once per point), Cc; write y1
            rows, n_, m_, ns_, c_ = a[:5]
            return rows * (4 + 4 * c_) + (rows // (m_ * ns_)) * n_ * c_ * 4 + (rows // ns_) * c_ * 4, 0
        if nm == "sg4d_group_sum_dy":            # (groups, ns, c1): read y1, dz1; write the per-centre sums
            return a[0] * a[2] * 4 * (2 * a[1] + 1), 0
        if nm in ("sg4d_ball_query_rows", "sg4d_ball_query_rows_indexed"):
            b_, n_, m_ = a[:3]
            return b_ * (12 * n_ + 12 * m_), 0
        if nm == "sg4d_linear_fwd":              # (rows, k, lda, n, group)
            return a[0] * 4 * (a[1] + a[3]), 2 * a[0] * a[1] * a[3]
        if nm == "sg4d_pool_bwd_da":             # (rows, C2, C1, group): read y2, y1; write dz1
            return a[0] * 4 * (a[1] + 2 * a[2]), 2 * a[0] * a[1] * a[2]
        if nm in ("sg4d_pool_bwd_dw", "sg4d_dense_pool_bwd_dw"):
            return a[0] * 4 * (a[1] + a[2]), 2 * a[0] * a[1] * a[2]
        if nm == "sg4d_inner_bwd_dx":            # (rows, C1, n, lddx, col0)
            return a[0] * 4 * (2 * a[1] + a[2]), 2 * a[0] * a[1] * a[2]
        if nm == "sg4d_inner_bwd_dw":            # (rows, C1, k, ldx)
            return a[0] * 4 * (2 * a[1] + a[2]), 2 * a[0] * a[1] * a[2]
        if nm == "sg4d_sa1_bwd_dw2_gram":        # Gram kernel + T1 kernel: each reads idx and the gathered points; T1 also dsel / garg
            rows, n_, m_, ns_, ps_, fs_ = a[:6]
            n2 = a[-1]
            cloud_bytes = min((rows // (m_ * ns_)) * n_, rows) * ps_ * 4
            return 2 * (rows * 4 + cloud_bytes) + (rows // ns_) * n2 * 5, 2 * rows * 64 * n2
        if nm in ("sg4d_sa1_fwd", "sg4d_sa1_bwd_da", "sg4d_sa1_bwd_dw2", "sg4d_sa_moments"):
            rows, n_, m_, ns_, ps_, fs_ = a[:6]       # then [foff,] c, n2 (zero-valued arguments are not recorded)
            n2 = a[-1] if nm != "sg4d_sa_moments" else 0
            clouds = rows // (m_ * ns_)
            cloud_bytes = min(clouds * n_, rows) * ps_ * 4      # the gathered points, each at most once
            pooled = (rows // ns_) * n2 * 5
            if nm == "sg4d_sa_moments":
                return rows * 4 + cloud_bytes, 0
            fl = 2 * rows * 64 * n2 + (2 * rows * 8 * 64 if nm != "sg4d_sa1_bwd_dw2" else 0)
            return rows * (4 + 4 * n2) + cloud_bytes + pooled, fl
        if nm == "sg4d_linear_fwd_grouped":      # (rows, n, m, ns, pstride, fstride, c, nout)
            rows, n_, m_, ns_, ps_, fs_ = a[:6]
            c_, nout = a[-2], a[-1]
            clouds = rows // (m_ * ns_)
            return rows * (4 + 4 * nout) + clouds * n_ * (fs_ + 3) * 4, 2 * rows * (c_ + 3) * nout
        if nm == "sg4d_inner_bwd_dw_grouped":    # (..., c, mout, lddw)
            rows, n_, m_, ns_, ps_, fs_ = a[:6]
            c_, mout = a[-3], a[-2]
            clouds = rows // (m_ * ns_)
            return rows * (4 + 8 * mout) + clouds * n_ * (fs_ + 3) * 4, 2 * rows * (c_ + 3) * mout
        if nm == "sg4d_dense_fwd":               # (rows, k, lda, n, ldy[, group])
            return a[0] * 4 * (a[1] + a[3]), 2 * a[0] * a[1] * a[3]
        if nm == "sg4d_dense_bwd_dx":            # (rows, kk, lda, nout, mode?, ...)
            return a[0] * 4 * (2 * a[1] + a[3]), 2 * a[0] * a[1] * a[3]
        if nm == "sg4d_dense_bwd_dw":            # (rows, m, lda, k, ...)
            return a[0] * 4 * (2 * a[1] + a[3]), 2 * a[0] * a[1] * a[3]
        return None, 0

    for k in kernels:
        try:
            bts, fl = model_of(k["call"], k["args"])
        except (IndexError, ZeroDivisionError):
            bts, fl = None, 0
        if bts:
            k["algorithmic_bytes"] = bts
            k["achieved_gbs"] = bts / (k["avg_ms"] * 1e-3) / 1e9
            k["hbm_frac"] = k["achieved_gbs"] / peak
        if fl:
            k["algorithmic_flops"] = fl
    # ---- roofline of the DOMINANT KERNEL: calls are grouped by the kernel that does their work; achieved =
    #      algorithmic bytes of all its launches / their summed duration (= bytes per launch / average launch duration)
    row_calls = ("sg4d_linear_fwd", "sg4d_pool_bwd_da", "sg4d_inner_bwd_dx", "sg4d_sa1_fwd", "sg4d_sa1_bwd_da",
                 "sg4d_linear_fwd_grouped", "sg4d_dense_fwd", "sg4d_dense_bwd_dx")
    wg_calls = ("sg4d_pool_bwd_dw", "sg4d_inner_bwd_dw", "sg4d_sa1_bwd_dw2", "sg4d_sa1_bwd_dw2_gram", "sg4d_inner_bwd_dw_grouped",
                "sg4d_dense_bwd_dw", "sg4d_dense_pool_bwd_dw")
    family = {c: "row_gemm_kernel" for c in row_calls}
    family.update({c: "wgrad_kernel" for c in wg_calls})
    family.update({"sg4d_fps_indexed": "fps_indexed_kernel", "sg4d_fps_rows": "fps_onchip_kernel",
                   "sg4d_ball_query_rows_indexed": "ball_query_kernel+ball_query_indexed_kernel",
                   "sg4d_ball_query_rows": "ball_query_kernel", "sg4d_group_rows": "group_rows_kernel",
                   "sg4d_group_rows_grad": "group_rows_grad_kernel", "sg4d_group_rows_grad_dy": "group_rows_grad_kernel",
                   "sg4d_gather_y1": "gather_y1_kernel", "sg4d_group_sum_dy": "group_sum_dy_kernel",
                   "sg4d_spatial_index_build": "spatial_build_kernel"})
    fam = {}
    for k in kernels:
        f = fam.setdefault(family.get(k["call"], k["call"]), {"ms": 0.0, "bytes": 0.0, "flops": 0.0, "launches": 0.0, "calls": set()})
        f["ms"] += k["ms_per_step"]
        f["bytes"] += k.get("algorithmic_bytes", 0) * k["launches_per_step"]
        f["flops"] += k.get("algorithmic_flops", 0) * k["launches_per_step"]
        f["launches"] += k["launches_per_step"]
        f["calls"].add(k["call"])

    traffic_file = None
    for cand in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        if os.path.exists(os.path.join(ROOT, "profiles", cand)):
            traffic_file = cand
            break

    def fam_roof(name, f):
        ach = f["bytes"] / (f["ms"] * 1e-3) / 1e9 if f["ms"] > 0 and f["bytes"] > 0 else None
        r = {"bound": "hbm", "kernel": name, "calls": sorted(f["calls"]), "launches_per_step": f["launches"],
             "avg_launch_ms": f["ms"] / max(f["launches"], 1e-9), "algorithmic_bytes_per_launch": f["bytes"] / max(f["launches"], 1e-9),
             "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if ach else None, "traffic": None,
             "peak_source": peak_src, "share_of_step": f["ms"] / table_ms_per_step}
        if traffic_file:   # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu capture
            t = json.load(open(os.path.join(ROOT, "profiles", traffic_file))).get("kernels", {}).get(name.split("+")[0])
            if t and t.get("launches"):
                r["traffic"] = t["dram_bytes"] / t["launches"]
                r["traffic_source"] = f"profiles/{traffic_file} (ncu, same workload; average per launch)"
        return r

    roof = None
    if fam:
        top_name = max(fam, key=lambda n: fam[n]["ms"])
        roof = fam_roof(top_name, fam[top_name])
        roof["others"] = {n: {"ms_per_step": f["ms"], "hbm_frac": (f["bytes"] / (f["ms"] * 1e-3) / 1e9 / peak) if f["bytes"] else None}
                          for n, f in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]) if n != top_name}
        # the shared MLP against the TENSOR roofline: algorithmic FLOPs of every layer evaluation of the
        # step / the summed duration of the two tensor-core kernels / the sustained dense bf16 peak (tf32 runs at half
        # of it, and 3xTF32 issues three products per algorithmic one: both are kept OUT of the numerator on purpose)
        mlp_ms = sum(fam[n]["ms"] for n in ("row_gemm_kernel", "wgrad_kernel") if n in fam)
        mlp_fl = sum(fam[n]["flops"] for n in ("row_gemm_kernel", "wgrad_kernel") if n in fam)
        if mlp_ms > 0:
            roof["mlp_tensor"] = {"bound": "tensor", "algorithmic_flops_per_step": mlp_fl, "ms_per_step": mlp_ms,
                                  "achieved": mlp_fl / (mlp_ms * 1e-3) / 1e12, "peak": tpeak, "unit": "TFLOP/s",
                                  "frac": mlp_fl / (mlp_ms * 1e-3) / 1e12 / tpeak,
                                  "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if pk else "fallback"}
        # the pair BASELINE.json's north star singles out: ball query + group.  The group is fused into the operand
        # stagers of the MLP kernels (the grouped tensor is never written), so the pair's time is the ball-query
        # kernels' and its bytes are the op-level contract of SURVEY.md 8(d): query (12n + 12m + 4 m ns) + grouped
        # output 4 (3 + C) m ns per scale and cloud.
        bq_ms, bq_bytes, calls = 0.0, 0.0, set()
        for k in kernels:
            if k["call"] in ("sg4d_ball_query_rows", "sg4d_ball_query_rows_indexed"):
                bq_ms += k["ms_per_step"]
                bq_bytes += k.get("algorithmic_bytes", 0) * k["launches_per_step"]
                calls.add(k["call"])
            if k["call"] in ("sg4d_sa1_fwd", "sg4d_linear_fwd_grouped", "sg4d_group_rows", "sg4d_gather_y1"):
                a = k["args"]
                # width of the grouped row the op produces: SA1 8 (padded xyz + features), materialised rows 3 + C, and for
                # the scales evaluated through the first layer's linearity the PROJECTED row (c1 channels) that is gathered
                width = {"sg4d_sa1_fwd": 8, "sg4d_linear_fwd_grouped": a[-2] + 3, "sg4d_gather_y1": a[4]}.get(k["call"], a[4] + 3)
                rows_ = a[0] if k["call"] != "sg4d_group_rows" else a[0] * a[2] * a[3]
                bq_bytes += rows_ * (4 + 4 * width) * k["launches_per_step"]
                if k["call"] in ("sg4d_group_rows", "sg4d_gather_y1"):   # group kernels that exist as launches of their own
                    bq_ms += k["ms_per_step"]
                calls.add(k["call"])
        if bq_ms > 0:
            roof["ball_query_plus_group"] = {"achieved": bq_bytes / (bq_ms * 1e-3) / 1e9, "unit": "GB/s",
                                             "frac": bq_bytes / (bq_ms * 1e-3) / 1e9 / peak, "ms_per_step": bq_ms,
                                             "share_of_step": bq_ms / table_ms_per_step, "calls": sorted(calls),
                                             "note": "time = the ball-query launches + the group kernels that run as launches "
                                                     "of their own (gather_y1, group_rows); bytes = query + grouped-output contract"}

    # ---- the REFERENCE ITSELF on this GPU: its own Python over its own kernels (oracle/ref_gpu.py), batch_size = 1 like
    #      its main.py, same synthetic scenes; fp32 with TF32 off, and fp16 autocast (its native `precision=16`)
    gpu_ref = None
    if world == 1 and not args.no_gpu_reference and not args.encoders_only:
        try:
            from oracle import ref_gpu
            if ref_gpu.available():
                del resident
                torch.cuda.empty_cache()
                rm = ref_gpu.build_model(dev)
                nsc = 1 if args.forward_only else 2
                scenes = [synthetic.to_device(synthetic.make_scene(900 + i, n_obj=args.n_obj, n_points_obj=args.points,
                                                                   n_points_rel=pr, pairs=args.pairs), dev) for i in range(nsc)]
                for sc in scenes:
                    sc["take_idx"] = 0
                v32, ms32, wall32 = ref_gpu.time_scenes(rm, scenes, reps=2)
                v16, ms16, wall16 = ref_gpu.time_scenes(rm, scenes, reps=2, autocast=True)
                gpu_ref = {"kind": "reference", "what": "the reference's own SGPNModelWrapper (Python, unmodified) over its own CUDA "
                           "kernels compiled for sm_100a, one scene per forward+backward (main.py's batch_size=1), same GPU",
                           "fp32_tf32_off": {"value": v32, "unit": "scenes/s", "ms_per_scene": ms32},
                           "fp16_autocast": {"value": v16, "unit": "scenes/s", "ms_per_scene": ms16},
                           "scenes_timed": 2 * nsc, "sg4d_over_reference_fp32": value / v32, "sg4d_over_reference_fp16": value / v16}
            else:
                gpu_ref = {"unavailable": "oracle/_ref/pn2_ref_ext.so or the staged reference Python is missing on this box"}
        except Exception as e:   # measurement infrastructure must never take the bench line down
            gpu_ref = {"unavailable": f"{type(e).__name__}: {e}"[:300]}

    cpu = None
    if not args.no_cpu_baseline:
        v, _, cpu = run_cpu(args, steps=2, warmup=1, budget_s=45.0)
        cpu["value"] = v

    line = {"metric": "scenes/sec fwd+bwd", "value": value, "unit": "scenes/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.bf16 else "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "scenes_per_gpu": S, "global_scenes": S * world,
                       "parallelism": f"dp{world} (scene-sharded, one grad all-reduce)",
                       "l2": f"inputs are {h2d_bytes / 1e6:.0f} MB per step per GPU (> 126 MB L2), no flush needed",
                       "peak_memory_gb": round(peak_mem_gb, 2),
                       "launch": "eager" if args.no_graph else "one CUDA-graph replay per step (capture of the eager step)",
                       "cpu_sample": cpu["sample"] if cpu else None},
            "clocks": clk.summary(), "e2e": e2e, "e2e_crops": e2e_crops, "gpu_launches": launches,
            "roofline": roof, "cpu_baseline": cpu, "gpu_reference": gpu_ref,
            "kernel_table": {"ms_per_step": table_ms_per_step, "own_kernel_ms_per_step": own_ms,
                             "note": "second pass of the same steps, one stream, CUDA events around every C-ABI call"},
            "kernels": kernels[:48]}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_sg4d(a)
