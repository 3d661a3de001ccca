#!/usr/bin/env python
"""Headline benchmark: forward+backward frames/s of the articulated-Gaussian-splat render path at 1920x1080 with a
500k-Gaussian composite hand+object scene (BASELINE.json), plus the HBM roofline of the dominant kernel, the end-to-end
number through the public API with host inputs, and the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = V training views per rank (--views-in-flight, default 4; gradient accumulation over the views of a step like the
reference's accum_iter), each: pose kernel (LBS + covariance + SH->RGB) -> rasterizer forward -> loss = sum(image * G) with a
fixed G ~ U[0,1] -> rasterizer backward -> pose backward accumulating into the flat per-Gaussian
gradient buffer; then (N > 1) one NCCL exchange of that buffer.  The V views of a step are independent until the
accumulation and run as parallel branches of one CUDA graph.  Views shard across ranks (weak scaling: every rank renders V
views per step); value = views (frames) per second over all ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fwd+bwd frames/sec at 1080p, 500k Gaussians; achieved HBM GB/s vs peak"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gaussians", type=int, default=500_000)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--views", type=int, default=50)
    ap.add_argument("--scene", default="composite", choices=["composite", "hand", "object"],
                    help="composite = BASELINE configs[3] (headline); hand = configs[2]; object = configs[1] (use --gaussians 100000 --width 800 --height 800)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--plain-allreduce", action="store_true",
                    help="N > 1: all-reduce the whole flat gradient buffer instead of the compact exchange (rank-one SH gradients)")
    ap.add_argument("--compact-exchange", action="store_true", help="use the compact exchange for any N > 1 (default: N <= 4)")
    ap.add_argument("--views-in-flight", type=int, default=4,
                    help="views per rank and step, captured on parallel streams of the step's CUDA graph and accumulated into one "
                         "gradient buffer (the reference's accum_iter); 1 = one view per step")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel from Python each step instead of replaying the captured CUDA graph")
    ap.add_argument("--chunks", type=int, default=-1,
                    help="N > 1: ranges of Gaussians whose pose backward + all-reduce are pipelined (manus_b200.dist.PipelinedStep); "
                         "0 = one all-reduce of the flat buffer after the step; -1 = auto (measured on B200 / NVSwitch: 0 at N = 2, 4 above)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "multimem", "p2p"],
                    help="N > 1: who sums the gradient ranges over the ranks -- NCCL (coalesced all-reduce per range) or the repository's "
                         "own multimem.ld_reduce / multimem.st kernel over NVSwitch multicast memory (csrc/exchange.cu); auto = multimem "
                         "for N >= 4 when the system offers multicast, the peer-to-peer kernel of the same file at N = 2 "
                         "(118 MB on two GPUs: p2p 238 us, NCCL 252 us, multimem 340 us)")
    ap.add_argument("--deferred-views", type=int, default=0,
                    help="N > 1: views per rank whose pose backward runs in the range-by-range tail (the others finish inside their branch); 0 = all")
    ap.add_argument("--exchange-ctas", type=int, default=32, help="CTAs of the multimem exchange kernel (it runs beside the pose backward)")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (BASELINE configs 1-3 and 5, drop-in and PyTorch-GPU baselines)")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every few milliseconds while the timed region runs (a timed
    region of the graph-replayed step lasts ~0.1 s, too short for nvidia-smi's polling); falls back to `nvidia-smi -lms`."""
    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, index: int, period_s: float = 0.004):
        self.index, self.period = index, period_s
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self.stop_flag, self.thread, self.proc, self.rows = False, None, None, []
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber devices: match by PCI bus id of the torch device
            import torch
            bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(torch.cuda.get_device_properties(index), "pci_bus_id") else None
            self.nv, self.h = pynvml, None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
                        self.h = h
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for name, bit in self.REASONS:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv is not None:
            self.stop_flag = False
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        else:
            q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                              "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.thread = threading.Thread(target=self._read, daemon=True)
                self.thread.start()
            except OSError:
                self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.sm.append(float(r[1]))
                self.max_mhz = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
                if len(r) > col and r[col].lower().startswith("active"):
                    self.reasons.add(name)

    def __exit__(self, *a):
        self.stop_flag = True
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        elif self.thread:
            self.thread.join(timeout=1)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(sm),
                "source": "nvml" if self.nv is not None else "nvidia-smi"}


def measured_traffic(kernel):
    """DRAM bytes (read + write) of one launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json,
    written by profiles/extract_traffic.py); None when the capture does not hold it."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        t = json.load(f).get(kernel)
    if not t or "dram_read_bytes" not in t:
        return None
    return t["dram_read_bytes"] + t.get("dram_write_bytes", 0.0)


def frame_bytes(n_hand, n_obj, D, P):
    """Algorithmic HBM bytes of one forward+backward frame (SURVEY.md section 8d / BASELINE.md section 4)."""
    return 1204 * n_hand + 1036 * n_obj + 244 * D + 40 * P


# algorithmic bytes of one launch of each kernel; N = Gaussians, D = instances, P = pixels
KERNEL_BYTES = {
    "pose_forward": lambda N, Nh, D, P, V: (236 + 52) * N + 84 * Nh,
    "pose_backward": lambda N, Nh, D, P, V: (52 + 236 + 236) * N + 84 * Nh,
    "preprocess": lambda N, Nh, D, P, V: 76 * N,
    "blend_forward": lambda N, Nh, D, P, V: 40 * D + 20 * P,
    "blend_backward": lambda N, Nh, D, P, V: 40 * D + 20 * P + 44 * V,
    "preprocess_backward": lambda N, Nh, D, P, V: 104 * N,
    "radix_scatter": lambda N, Nh, D, P, V: None, "radix_hist": lambda N, Nh, D, P, V: None,
}


def cpu_baseline_frame(scene, width, height, view, threads):
    """One forward+backward frame of the same workload on the host: PyTorch-CPU restatement of the reference's pose step
    (oracle/pose_ref.py, autograd) + the C restatement of the rasterizer (oracle/raster_ref.c) on `threads` threads."""
    import numpy as np
    import torch

    from manus_b200 import synth
    from oracle import pose_ref
    from oracle.raster_ref import RasterRef

    torch.set_num_threads(threads)
    cam = synth.camera(view, width, height)
    t = lambda a: torch.tensor(a)
    names = ["xyz", "log_scale", "quat", "opacity_logit", "f_dc", "f_rest"]
    G = np.random.default_rng(7).uniform(0, 1, (3, height, width)).astype(np.float32)
    ref = RasterRef("f32")
    t0 = time.perf_counter()
    leaves = {k: t(getattr(scene, k)).requires_grad_(True) for k in names}
    nh = scene.n_hand
    parts = []
    if nh:
        tfs = pose_ref.bone_transforms(t(synth.posed_bones(view)), t(scene.bones_rest), True)
        parts.append(pose_ref.pose_gaussians_ref(*[leaves[k][:nh] for k in names], t(scene.skin_wts), tfs, t(cam.camera_center)))
    if nh < scene.n:
        parts.append(pose_ref.pose_gaussians_ref(*[leaves[k][nh:] for k in names], None, None, t(cam.camera_center)))
    outs = [torch.cat([p[i] for p in parts], 0) for i in range(4)]
    t1 = time.perf_counter()
    img, radii, D = ref.forward(outs[0].detach().numpy(), outs[3].detach().numpy(), colors_precomp=outs[2].detach().numpy(),
                                cov3D_precomp=outs[1].detach().numpy(), viewmatrix=cam.world_view_transform,
                                projmatrix=cam.full_proj_transform, campos=cam.camera_center, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
                                W=width, H=height, bg=np.ones(3, np.float32), nthreads=threads)
    t2 = time.perf_counter()
    g = ref.backward(G, nthreads=threads)
    t3 = time.perf_counter()
    torch.autograd.backward(outs, [t(g["means3D"]), t(g["cov3D"]), t(g["colors"]), t(g["opacity"])])
    t4 = time.perf_counter()
    ref.close()
    return dict(total_s=t4 - t0, pose_fwd_s=t1 - t0, raster_fwd_s=t2 - t1, raster_bwd_s=t3 - t2, pose_bwd_s=t4 - t3, D=D)


def run_reference(args, rank, world):
    """Reference arm: the reference's path on the host cores.  The reference's own CUDA rasterizer is not part of
    /root/reference (third-party, cloned at install time) so nothing can be compiled into oracle/_ref; this times the CPU
    port (oracle/) with all host threads, one full frame per step."""
    if rank != 0:
        return
    from manus_b200 import synth

    threads = os.cpu_count() or 1
    scene = make_scene(args)
    budget_s = 150.0          # the whole reference run must end within a few minutes
    t_warm = []
    for i in range(min(max(args.warmup, 1), 1)):
        t_warm.append(cpu_baseline_frame(scene, args.width, args.height, i, threads)["total_s"])
    # a step = one full frame on the host; when K frames do not fit the budget, time as many as fit
    k_eff = max(1, min(args.steps, int(budget_s / max(t_warm[-1], 1e-3))))
    times = []
    for i in range(k_eff):
        times.append(cpu_baseline_frame(scene, args.width, args.height, i % args.views, threads)["total_s"])
    ms = 1e3 * sum(times) / len(times)
    val = 1e3 / ms
    sample = (f"one full {args.width}x{args.height} frame per step (pose fwd+bwd in PyTorch-CPU, raster fwd+bwd in C, {threads} threads); "
              f"{k_eff} of the requested {args.steps} steps timed to stay within {budget_s:.0f} s")
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": k_eff, "warmup": len(t_warm),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference",
            "config": workload_config(args, 1, max(1, args.views_in_flight)),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local_rank):
    """Pinned host buffers are placed on the NUMA node of the thread that first touches them: run this rank on the cores of the
    node its GPU hangs off, so that the host->device copies of the ranks do not all read one socket's memory.  Best effort."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev_id = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return {"unavailable": f"{path} reports no NUMA node"}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return {"numa_node": node, "cpus": len(allowed)}
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"}
    return {"unavailable": "no allowed CPU on the GPU's NUMA node"}


def make_scene(args):
    from manus_b200 import synth

    if args.scene == "hand":
        return synth.make_hand(args.gaussians, seed=0)
    if args.scene == "object":
        return synth.make_object(args.gaussians, seed=1)
    return synth.make_composite(args.gaussians, seed=0)


def workload_config(args, world, vif=1):
    kind = {"composite": "composite hand+object {n} Gaussians (60% skinned by 20+1 bones, 40% static)",
            "hand": "articulated hand {n} Gaussians (all skinned by 20+1 bones)", "object": "static object {n} Gaussians (no skinning)"}[args.scene]
    return {"workload": kind.format(n=args.gaussians) + ", " +
                        f"{args.views} shipped views/poses at {args.width}x{args.height}, SH degree 3, white background",
            "global_views_per_step": world * vif, "views_in_flight_per_gpu": vif,
            "parallelism": f"view-sharded dp{world}, {vif} view(s) per rank and step accumulated into one gradient buffer (parallel branches "
                           "of one CUDA graph), one exchange of per-Gaussian gradients per step",
            "l2": "working set per step (parameters 118 MB + gradients 118 MB + instance records) exceeds the 126 MB L2 and the view "
                  "changes every step; no explicit flush",
            "loss": "sum(image * G), G ~ U[0,1] fixed (seed 7): one dot product; its gradient G is the seed of the backward"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    from manus_b200 import _lib, synth
    from manus_b200 import build as mb_build
    from manus_b200.dist import SceneRenderer
    from manus_b200.rasterizer import set_capacity_mode

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: manus_b200 has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)      # before any pinned allocation
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        mb_build.build()
    if world > 1:
        dist.barrier()
    _lib.lib()

    W, H, K, WU = args.width, args.height, args.steps, max(args.warmup, 3)
    scene = make_scene(args)
    r = SceneRenderer(scene, dev, W, H)
    mc_exchange, mc_note = None, None
    if args.chunks < 0:
        args.chunks = 4 if world >= 4 else 0
    if args.exchange == "auto":
        args.exchange = "multimem" if world >= 4 else "p2p"
    if world > 1 and args.exchange in ("multimem", "p2p"):
        # the gradient buffer moves into NVSwitch multicast memory BEFORE anything captures its address
        try:
            from manus_b200.exchange import MulticastExchange
            mc_exchange = MulticastExchange(r.flat)
        except Exception as e:      # no multicast on this system: NCCL does the exchange (said in the line)
            mc_note = f"multimem unavailable ({type(e).__name__}: {e}); NCCL used"
    n_hand, n_obj = scene.n_hand, scene.n - scene.n_hand
    views = list(range(args.views))
    VIF = max(1, args.views_in_flight)      # with --no-graph the views of a step are enqueued one after the other
    my_view = lambda it, slot=0: views[((it * VIF + slot) * world + rank) % len(views)]
    # target / loss weights G ~ U[0,1]: 8-bit like the dataset's images (rgb / 255, src/datasets/brics_dynamic.py), so that the
    # end-to-end step ships one byte per channel over PCIe and converts on the device; the resident steps keep the fp32 copy
    # (stored [3,H,W] like the rasterizer's output and viewed as [H,W,3]: image and target then flatten without a copy)
    G_u8_host = torch.randint(0, 256, (3, H, W), generator=torch.Generator().manual_seed(7), dtype=torch.uint8).pin_memory().permute(1, 2, 0)
    G_u8_dev = torch.empty((3, H, W), dtype=torch.uint8, device=dev).permute(1, 2, 0)
    G_u8_dev.copy_(G_u8_host)
    G_dev = G_u8_dev.float() / 255.0                  # keeps the [H,W,3]-view-of-[3,H,W] layout
    G_u8_host_hwc = G_u8_host.contiguous().pin_memory()      # dense HWC copy for the reference's loss kernel (reads gt row by row)
    staged = {}
    for v in views:
        _, c, b = r.view_inputs_host(v)
        staged[v] = (c.to(dev), b.to(dev))

    # ---- untimed: instance / visible counts of every view (exact mode), then reserve mode (no host read-back per frame)
    from manus_b200 import rasterizer as _rz
    from manus_b200.dist import CAM_FLOATS, GraphedStep
    set_capacity_mode("exact")
    D_all, V_all = view_counts(r, staged, views)
    set_capacity_mode("reserve", margin=1.1)
    _rz.reserve_capacity(dev.index, scene.n, H, W, max(D_all.values()))

    # The probe loss of SURVEY.md section 8d, sum(image * G): one dot product forward, and its gradient IS G -- handed to the backward
    # as the seed instead of letting autograd form it (the loss is there to time the path, not itself).  image is the permuted
    # [H,W,3] view of the rasterizer's [3,H,W] output; G is stored the same way, so both flatten without a copy.
    def loss_fn(image, target):
        return torch.dot(image.permute(2, 0, 1).reshape(-1), target.permute(2, 0, 1).reshape(-1)), target

    # N > 1: the backward skips the f_rest gradient and the ranks exchange the rank-one SH gradient factors instead of
    # all-reducing all 59 floats per Gaussian (manus_b200.dist.CompactGradExchange)
    # measured on B200 / NVSwitch: compact wins at N = 2 (0.838 vs 0.948 ms/step); at N = 8 the rebuild over 8 views costs
    # what the smaller collective saves (1.060 vs 1.066 ms/step) and the extra launches hurt the per-step-synchronised e2e
    # loop, so the default switches to the plain all-reduce above 4 ranks
    # the compact exchange carries ONE view per rank; with several views per rank and step the flat buffer is all-reduced
    compact = world > 1 and VIF == 1 and args.compact_exchange and not args.plain_allreduce
    # the step object: PipelinedStep whenever a step holds several views (their pose backwards run as ONE multi-view pass over the
    # Gaussians, range by range when there is an exchange to overlap); GraphedStep for one view per step
    pipelined = not compact and not args.no_graph and (VIF > 1 or (world > 1 and not args.plain_allreduce and args.chunks > 0))
    n_chunks = args.chunks if (world > 1 and args.chunks > 0 and not args.plain_allreduce) else 1
    from manus_b200.dist import CompactGradExchange, PipelinedStep
    exchange = CompactGradExchange(r) if compact else None

    def make_step(loss, target_like, vif=VIF, stats=None):
        if args.no_graph:
            return None
        if pipelined and (vif > 1 or world > 1):
            # the pose backward of all the step's views runs range by range over the Gaussians; with N > 1 each finished range is
            # summed over the ranks on a side stream while the next one computes
            return PipelinedStep(r, loss, target_like, view=views[0], views_in_flight=vif, chunks=n_chunks, stats=stats,
                                 exchange=mc_exchange if n_chunks > 1 else None, exchange_ctas=args.exchange_ctas,
                                 deferred_views=args.deferred_views or None, reduce=n_chunks > 1)
        return GraphedStep(r, loss, target_like, view=views[0], compact_sh=compact, views_in_flight=vif, stats=stats)

    graphed = make_step(loss_fn, G_dev)

    def reduce_gradients():
        if pipelined and n_chunks > 1:
            return                       # inside PipelinedStep.replay
        if exchange is not None:
            exchange()
        elif world > 1 and mc_exchange is not None:
            mc_exchange.all_reduce_all(args.exchange_ctas, p2p=args.exchange == "p2p")
        elif world > 1:
            dist.all_reduce(r.flat.grad)

    def step_resident(it, eager=False, graphed=graphed):
        if graphed is not None and not eager:
            # the whole step (per view: pose fwd -> raster fwd -> loss -> raster bwd -> pose bwd) is ONE graph launch; the
            # per-view camera / bones are copied device-to-device into the graph's static inputs
            for slot in range(graphed.V):
                v = my_view(it, slot)
                graphed.set_inputs(staged[v][0], staged[v][1], None, slot=slot)
            loss = graphed.replay()
        else:
            loss = 0.0
            for slot in range(VIF):
                v = my_view(it, slot)
                out = r.render(v, sink=r.flat.grads, cam_dev=staged[v][0], bones_dev=staged[v][1], compact_sh=compact, accumulate=slot > 0)
                l, seed = loss_fn(out["render"], G_dev)
                out["render"].backward(seed)
                loss = loss + l.detach()
        reduce_gradients()
        return loss

    # End-to-end step: the step's inputs (per view: packed camera 39 floats, posed bones 320 floats, target image H*W*3 floats)
    # come from pinned host memory every step and the step's loss is read back every step.  Two input slots: the copies of
    # step i+1 are enqueued on a copy stream while step i computes, so the PCIe transfer overlaps the kernels; every copy is
    # inside the timed region.  With graph replay each slot is the static input set of its own captured graph (the host
    # copies land where the kernels read: no device-to-device staging).  The loss goes to pinned host memory with an
    # asynchronous copy and is read by the host one step later (before step i+1 is enqueued the host waits for the loss of
    # step i-1), so the GPU queue never runs dry; the last timed step waits for its own loss inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    NSLOT = 3      # input sets in flight: the copies of steps i+1 and i+2 are enqueued while step i computes

    class E2E:
        """loss_fn(image, target_u8): the target arrives as uint8 [H,W,3] (one byte per channel over PCIe)."""

        def __init__(self, loss_fn, host_target=None):
            self.loss_fn = loss_fn
            self.host_target = G_u8_host if host_target is None else host_target
            like = G_u8_dev if host_target is None else host_target.to(dev)
            self.graphs = None if args.no_graph else [make_step(loss_fn, like) for _ in range(NSLOT)]
            self.slots = []
            for k in range(NSLOT):
                if self.graphs is not None:
                    g = self.graphs[k]
                    d = dict(g=g.targets, cam=g.cams, bones=g.bones_all)
                else:
                    d = dict(g=[torch.empty_like(like) for _ in range(VIF)], cam=[torch.empty(CAM_FLOATS, device=dev) for _ in range(VIF)],
                             bones=[torch.empty(320, device=dev) for _ in range(VIF)])
                d.update(ready=torch.cuda.Event(), free=torch.cuda.Event(), loss_host=torch.zeros(1).pin_memory(),
                         loss_done=torch.cuda.Event(), staged=None, pending=False)
                self.slots.append(d)
            self.last_loss = None
            self.final_it = -1

        def stage(self, it):
            slot = self.slots[it % NSLOT]
            if slot["staged"] == it:
                return
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(slot["free"])           # the step that last used this slot has finished with it
                for j in range(VIF):                           # every view of the step: target image, camera, posed bones
                    _, c, b = r.view_inputs_host(my_view(it, j))
                    slot["g"][j].copy_(self.host_target, non_blocking=True)
                    slot["cam"][j].copy_(c, non_blocking=True)
                    slot["bones"][j].copy_(b, non_blocking=True)
                slot["ready"].record(copy_stream)
            slot["staged"] = it

        def collect(self, slot):
            if slot["pending"]:
                slot["loss_done"].synchronize()
                self.last_loss = float(slot["loss_host"])       # the step's result on the host
                slot["pending"] = False

        def __call__(self, it):
            slot = self.slots[it % NSLOT]
            self.stage(it)                                      # first step of a run: nothing was prefetched
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(slot["ready"])
            if self.graphs is not None:
                loss = self.graphs[it % NSLOT].replay()
            else:
                loss = 0.0
                for j in range(VIF):
                    out = r.render(my_view(it, j), sink=r.flat.grads, cam_dev=slot["cam"][j], bones_dev=slot["bones"][j],
                                   compact_sh=compact, accumulate=j > 0)
                    res = self.loss_fn(out["render"], slot["g"][j])
                    if isinstance(res, tuple):
                        out["render"].backward(res[1])
                        res = res[0]
                    else:
                        res.backward()
                    loss = loss + res.detach()
            reduce_gradients()
            slot["loss_host"].copy_(loss.reshape(1), non_blocking=True)     # device -> host read of the step's result
            slot["loss_done"].record(cur)
            slot["pending"] = True
            slot["free"].record(cur)
            # the slot that step it + NSLOT - 1 will use held step it - 1: read that step's loss, then refill the slot
            self.collect(self.slots[(it + NSLOT - 1) % NSLOT])
            for ahead in range(1, NSLOT):
                self.stage(it + ahead)                          # host->device copies of the next steps (copy stream)
            if it == self.final_it:
                for sl in self.slots:
                    self.collect(sl)
            return self.last_loss

        def check(self):
            if self.graphs is not None:
                for g in self.graphs:
                    g.check()

    def timed(fn, steps, sampler=None):
        for it in range(WU):
            fn(it)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if hasattr(fn, "final_it"):
            fn.final_it = WU + steps - 1        # the last timed step reads its own result before the closing event
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx = sampler if sampler is not None else _Null()
        with ctx:
            e0.record()
            for it in range(steps):
                fn(WU + it)
            e1.record()
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    # host time to enqueue one step (no synchronisation inside): what bounds a step when the GPU is faster than Python
    for it in range(WU):                 # first calls pay one-time costs (stream / event creation, the exchange's first barrier)
        step_resident(it)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(10):
        step_resident(it)
    host_enqueue_ms = (time.perf_counter() - t0) / 10 * 1e3
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    ms_step = timed(step_resident, K, sampler)
    clocks = sampler.summary()
    # the same loss on a uint8 target: sum(image * g) / 255 (type promotion inside the one elementwise kernel)
    def loss_fn_u8(image, target):
        g = target * (1.0 / 255.0)                    # one elementwise kernel: uint8 -> fp32, same layout
        return torch.dot(image.permute(2, 0, 1).reshape(-1), g.permute(2, 0, 1).reshape(-1)), g

    e2e_step = E2E(loss_fn_u8)
    ms_e2e = timed(e2e_step, K)
    e2e_step.check()
    del e2e_step
    # the same end-to-end step with the reference's training loss 0.8 L1 + 0.2 (1 - SSIM) (fused kernel, manus_b200.losses)
    from manus_b200.losses import photometric_loss
    e2e_photo = E2E(lambda image, target: photometric_loss(image, target * (1.0 / 255.0), 0.8, 0.2), host_target=G_u8_host_hwc)
    ms_e2e_photo = timed(e2e_photo, K)
    e2e_photo.check()
    del e2e_photo
    # one view per step (the latency of a single frame; what earlier revisions of this bench reported as `value`)
    ms_single = None
    if graphed is not None and VIF > 1 and world == 1:
        graphed_one = make_step(loss_fn, G_dev, vif=1)
        ms_single = timed(lambda it: step_resident(it, graphed=graphed_one), K)
        graphed_one.check()
        del graphed_one
    # the step with the densification statistics of the reference's density_update (gaussian.py:335-338, gaussian_utils.py:461-473:
    # every step while global_step < densify_until_step) updated inside the pose backward kernel, and their reduction over the
    # ranks (SUM / SUM / MAX; once per densification interval of 100 steps, not per step)
    densify = None
    if graphed is not None:
        from manus_b200.densify import GaussianState
        gs = GaussianState(r.flat)
        stats_step = make_step(loss_fn, G_dev, stats=(gs.xyz_gradient_accum, gs.denom, gs.max_radii2D))
        ms_stats = timed(lambda it: step_resident(it, graphed=stats_step), max(20, K // 4))
        stats_step.check()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gs.reduce_stats()
        e1.record()
        torch.cuda.synchronize()
        densify = {"ms_per_step_with_stats": ms_stats, "ms_per_step": ms_step, "reduce_stats_ms": e0.elapsed_time(e1),
                   "reduce_every_steps": 100, "visible_updates_seen": float(gs.denom.max()),
                   "note": "statistics updated by the pose backward kernel (atomics; V views in flight); reduce_stats all-reduces "
                           "3 x N floats once per densification interval (config/model/gaussian/gaussian.yaml:15)"}
        del stats_step
    # N > 1: the exchanged gradients against rank 0 rendering ALL the step's views alone with gradient accumulation
    # (hand_dynamic.py:248,259-277: the R-rank step must equal the 1-rank step with accum_iter = R x V)
    grad_check = None
    if world > 1 and graphed is not None:
        for slot in range(graphed.V):
            v = my_view(0, slot)
            graphed.set_inputs(staged[v][0], staged[v][1], None, slot=slot)
        graphed.replay()
        reduce_gradients()
        torch.cuda.synchronize()
        got = r.flat.grad.clone()
        if rank == 0:
            want = torch.zeros_like(got)
            for rk in range(world):
                for slot in range(VIF):
                    v = views[((0 * VIF + slot) * world + rk) % len(views)]
                    out = r.render(v, sink=r.flat.grads, cam_dev=staged[v][0], bones_dev=staged[v][1])
                    out["render"].backward(loss_fn(out["render"], G_dev)[1])
                    want += r.flat.grad
            torch.cuda.synchronize()
            scale = float(want.abs().max())
            grad_check = {"max_abs_err": float((got - want).abs().max()), "max_abs_grad": scale,
                          "max_rel_err": float((got - want).abs().max()) / max(scale, 1e-30),
                          "rel_l2_err": float((got - want).double().norm() / want.double().norm().clamp_min(1e-30)),
                          "views": world * VIF, "note": "all-reduced flat gradient buffer of one step vs the same views accumulated on rank 0"}
        dist.barrier()
    # reserve mode reads nothing back per frame: make sure no timed frame ran out of instance capacity
    if graphed is not None:
        graphed.check()
    _rz.check_overflow()

    # ---- per-kernel pass (CUDA events around every launch, on the launching stream): roofline of the dominant kernel
    _lib.profile_enable(True)
    _lib.profile_report()
    for it in range(K):
        step_resident(WU + it, eager=True)     # same kernels as the graph, launched one by one so that each can be timed
    torch.cuda.synchronize()
    prof = _lib.profile_report()
    _lib.profile_enable(False)
    seen = [my_view(WU + it, j) for it in range(K) for j in range(VIF)]
    D_mean = float(np.mean([D_all[v] for v in seen]))
    V_mean = float(np.mean([V_all[v] for v in seen]))
    P = W * H
    peak, peak_src = measured_peaks()
    total_ms = sum(ms for _, ms in prof.values())
    top = max(prof.items(), key=lambda kv: kv[1][1])
    top_name, (top_n, top_ms) = top
    per_launch_ms = top_ms / top_n
    fb = KERNEL_BYTES.get(top_name, lambda *a: None)(scene.n, n_hand, D_mean, P, V_mean)
    roofline = {"bound": "hbm", "kernel": top_name, "achieved": (fb / (per_launch_ms * 1e-3) / 1e9) if fb else None, "peak": peak,
                "unit": "GB/s", "frac": (fb / (per_launch_ms * 1e-3) / 1e9 / peak) if fb else None, "traffic": measured_traffic(top_name),
                "traffic_source": "profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture, per launch)",
                "peak_source": peak_src, "launch_ms": per_launch_ms, "share_of_step": top_ms / total_ms if total_ms else None,
                "algorithmic_bytes_per_launch": fb,
                "timing": "CUDA events around every launch on the launching stream, kernels enqueued one by one (no overlap between views)",
                "kernels_ms_per_frame": {k: round(ms / (K * VIF), 5) for k, (n, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])}}
    launches_per_step = sum(n for n, _ in prof.values()) / K
    fbytes = frame_bytes(n_hand, n_obj, D_mean, P)
    value = world * VIF * 1e3 / ms_step
    e2e_val = world * VIF * 1e3 / ms_e2e
    h2d = VIF * (G_u8_host.numel() + (CAM_FLOATS + 320) * 4)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": WU, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world, VIF), "clocks": clocks, "frames_per_step": world * VIF,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e},
            "e2e_reference_loss": {"value": world * VIF * 1e3 / ms_e2e_photo, "unit": UNIT, "ms_per_step": ms_e2e_photo,
                                   "loss": "0.8 * L1 + 0.2 * (1 - SSIM) as in config/COMPOSITE.yaml:22-23 (fused kernel), host inputs as in e2e"},
            "gpu_launches": int(round(launches_per_step * K)), "host_enqueue_ms_per_step": host_enqueue_ms,
            "launch_mode": "eager (one launch per kernel)" if graphed is None else "CUDA graph replay (one launch per step)",
            "single_view": None if ms_single is None else {"value": 1e3 / ms_single, "unit": UNIT, "ms_per_step": ms_single,
                                                           "note": "one view per step (views_in_flight = 1), inputs resident"},
            "roofline": roofline,
            "frame": {"num_rendered_mean": D_mean, "visible_mean": V_mean, "algorithmic_bytes": fbytes,
                      "achieved_gbps": fbytes * (VIF * 1e3 / ms_step) / 1e9, "frac_of_hbm_peak": fbytes * (VIF * 1e3 / ms_step) / 1e9 / peak,
                      "allreduce_bytes": r.flat.allreduce_bytes() if world > 1 else 0,
                      "exchange": ("none" if world == 1 else f"pipelined: multi-view pose backward over {n_chunks} ranges of Gaussians, each range's six "
                                   "gradient pieces summed over the ranks on a side stream while the next range computes" if (pipelined and n_chunks > 1)
                                   else "compact: all-gather of the DC gradients (12 B per Gaussian and rank) + all-reduce of "
                                   "the 11 non-SH floats + local rebuild of the SH gradients" if compact else "one sum of the flat gradient buffer over the ranks after the step"),
                      "exchange_bytes_per_rank": (0 if world == 1 else (scene.n * (12 * world + 44)) if compact else r.flat.allreduce_bytes())}}
    line["exchange_impl"] = (None if world == 1 else mc_note or "NCCL all-reduce" if mc_exchange is None else
                             "peer-to-peer load / add / store kernel over NVLink (csrc/exchange.cu: mb_p2p_allreduce)" if args.exchange == "p2p" and n_chunks == 1
                             else "multimem.ld_reduce / multimem.st kernel over NVSwitch multicast memory (csrc/exchange.cu)")
    line["densification_stats"] = densify
    line["grad_check"] = grad_check
    line["numa_binding"] = numa
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cpu_baseline_frame(scene, W, H, 0, threads)                    # warm-up (thread pools, page faults, library load)
        cbs = sorted((cpu_baseline_frame(scene, W, H, v, threads) for v in (0, 1, 2)), key=lambda c: c["total_s"])
        cb = cbs[1]                                                    # median of 3
        line["cpu_baseline"] = {"value": 1.0 / cb["total_s"], "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "full 1080p frames of the same scene (views 0-2; one warm-up frame, median of 3): pose fwd+bwd in "
                                          "PyTorch-CPU + raster fwd+bwd in C (oracle/), all host threads",
                                "breakdown_s": {k: round(v, 4) for k, v in cb.items() if k != "D"}}
    if rank == 0 and world == 1 and not args.no_extras and graphed is not None:
        import bench_extras as bx
        torch.cuda.empty_cache()
        ex = {}
        try:
            ex["config1_pose_only_50k"] = bx.pose_only(dev)
            ex["config2_object_100k_800x800"] = bx.config_run(dev, "object", 100_000, 800, 800, 1, 1, 100, peak, seed=1)
            ex["config3_hand_300k_1080p_50views"] = bx.config_run(dev, "hand", 300_000, 1920, 1080, 50, VIF, 40, peak)
            ex["config5_sweep_1gpu"] = [bx.config_run(dev, "composite", n, 1920, 1080, 1, VIF, 16, peak) for n in
                                        (50_000, 100_000, 200_000, 500_000, 1_000_000, 2_000_000)]
            ex["dropin"] = bx.dropin_lines(dev, scene, W, H, views)
            up = bx.upstream_rasterizer_line(dev, scene, W, H, views)
            if up is not None:
                line["reference_cuda_rasterizer"] = up
        except Exception as e:          # a secondary measurement must not cost the headline line
            ex["error"] = f"{type(e).__name__}: {e}"
        line["extras"] = ex
    # north_star asks for the reference's own CUDA rasterizer on one GPU next to this number: it is a third-party submodule
    # cloned at install time (setup_env.sh:6-13), absent from the reference tree and from this image (no network)
    line["reference_cuda_rasterizer"] = {"value": None, "unavailable": "diff-gaussian-rasterization / simple-knn sources are not part of "
                                         "the reference checkout and cannot be fetched here; the reference arm is the CPU port (cpu_baseline)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def view_counts(r, staged, views):
    """(num_rendered, num_visible) of every view via mb_raster_query, for the bytes-per-frame bookkeeping."""
    import torch

    from manus_b200 import rasterizer as rz
    from manus_b200.pose import bone_transforms, pose_gaussians

    Ds, Vs = {}, {}
    for v in views:
        cam, _, _ = r.view_inputs_host(v)
        with torch.no_grad():
            leaves = [p.detach() for p in r.flat.leaves()]
            bone_tf = bone_transforms(staged[v][1].view(-1, 4, 4), r.bones_rest, True) if r.n_hand else None
            px, pc, col, op = pose_gaussians(*leaves, r.skin, bone_tf, staged[v][0][32:35], r.sh_degree, r.flat.isotropic, r.n_hand)
            settings = rz.GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, r.bg, 1.0, staged[v][0][0:16],
                                                        staged[v][0][16:32], r.sh_degree, staged[v][0][32:35], False, False)
            _, _, st = rz.rasterize_forward(settings, px, op.reshape(-1), colors_precomp=col, cov3D_precomp=pc)
            Ds[v], Vs[v], _ = rz.raster_query(st)
    return Ds, Vs


if __name__ == "__main__":
    main()
