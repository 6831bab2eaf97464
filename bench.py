#!/usr/bin/env python
"""bench.py -- person-crops/s of the I2R-Net forward (BASELINE config C2: vanilla HRNet-W48-S,
256x192, 8 images x 4 persons = 32 crops per GPU per step), one process per GPU.

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...     # the reference algorithm on the host CPU cores (oracle port)

One JSON line on stdout (rank 0).  `value` = crops/s with inputs resident in HBM; `e2e` = the same
through the public module call with pinned HOST inputs and a host read of the heatmaps every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

import paths  # noqa: F401

GFLOP_PER_CROP = {"C2": 19.43, "C3": 43.75, "C4": 28.24, "C5": 62.84}      # algorithmic, 2*MAC, BASELINE.md section 2
# BASELINE.json configs that fit one GPU: C2 is the configuration the metric is quoted on (the default workload);
# C3 (TransPose-H two-stage, split-operand precision) is selectable for measurements of that family.
WORKLOADS = {
    "C2": dict(yaml="coco/interformer_coco_w48_pure_en6.yaml", images=8, persons=4,
               text="C2: vanilla I2R-Net (interformer_pureMulti) HRNet-W48-S 256x192, 8 images x 4 persons = 32 crops "
                    "per GPU per step, whole images per rank (no data-path collective)",
               precision="fp16 operands, fp32 accumulate (tcgen05 kind::f16), single pass", dtype="f16"),
    "C3": dict(yaml="coco/interformer_coco_tph_192_p4_b4.yaml", images=4, persons=4,
               text="C3: two-stage I2R-Net (interformer_2stage), TransPose-H first stage (6 intra layers over 3072 "
                    "tokens per crop) + 4 inter layers, 256x192, 4 images x 4 persons = 16 crops per GPU per step",
               precision="split-operand fp16 pairs (hi+lo), three-term products, fp32 accumulate (tcgen05 kind::f16)",
               dtype="f16x2"),
    "C4": dict(yaml="coco/interformer_coco_hrt_192_p2_b12.yaml", images=1, persons=8,
               text="C4: HRFormer-B + I2R-Net (interformer), 256x192, 1 image x 8 persons = 8 crops per GPU per step "
                    "(BASELINE's 64 crops over 8 GPUs = one image per rank; first correct path: window attention on "
                    "mma.sync, unfused blocks)",
               precision="split-operand fp16 pairs (hi+lo), three-term products, fp32 accumulate (tcgen05 kind::f16); "
                         "channels 78/156/312/624 zero-padded to multiples of 16",
               dtype="f16x2"),
    "C5": dict(yaml="coco/interformer_coco_hrt_288_p2_b4.yaml", images=1, persons=12, hw=(384, 288),
               text="C5: HRFormer-B + I2R-Net (interformer), 384x288, 1 image x 12 persons = 12 crops per GPU per step "
                    "(BASELINE's 96 crops over 8 GPUs = one image per rank; inter-human sequences of 5184 tokens)",
               precision="split-operand fp16 pairs (hi+lo), three-term products, fp32 accumulate (tcgen05 kind::f16); "
                         "channels 78/156/312/624 zero-padded to multiples of 16",
               dtype="f16x2"),
}


def _smem_port(smem_bytes, launch_ms, clocks, dev):
    """Second roofline of the dominant launch class: shared-memory port.  At the output widths of this model (48 / 96
    channels per MMA) the tensor core's own operand fetch -- A 4 KB + B 32*N bytes per M128 x N x K16 MMA, re-read per
    tap -- saturates the 128 B/clk shared-memory port of an SM before the tensor pipe is busy, so the algorithmic
    shared-memory bytes of a launch (i2r_b200.ops.halo_smem_bytes) over its measured duration is the fraction of the
    BINDING resource in use; it counts start-up, tile-count quantisation and the tail of the launch as idle."""
    import torch
    if not smem_bytes or launch_ms <= 0:
        return None
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    mhz = (clocks or {}).get("sm_mhz") or 1965
    peak = 128.0 * sms * mhz * 1e6 / 1e12          # TB/s
    achieved = smem_bytes / (launch_ms * 1e-3) / 1e12
    return {"achieved": achieved, "peak": peak, "unit": "TB/s", "frac": achieved / peak,
            "bytes_per_launch": smem_bytes, "peak_kind": "128 B/clk/SM x %d SMs x %d MHz (sampled)" % (sms, mhz)}


def _ncu_traffic(kernel_substr):
    """DRAM read + write bytes per launch of `kernel_substr` from the newest committed `ncu --set full` summary
    (tools/ncu_summary.py output under profiles/), or None.  Parsed, not a literal."""
    import glob
    import re
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for path in sorted(glob.glob(os.path.join(paths.REPO, "profiles", "r*_ncu_*full_summary.txt")), reverse=True):
        rows = {}
        with open(path) as f:
            for line in f:
                m = re.match(r"(\S[^\[]*?)\s{2,}(.*?)\s+\[(.*?)\]\s*$", line)
                if m:
                    rows[m.group(1).strip()] = ([v.strip() for v in m.group(2).split("|")], m.group(3))
        names = rows.get("Kernel Name", ([], ""))[0]
        cols = [i for i, n in enumerate(names) if kernel_substr in n]
        if not cols or "dram__bytes_read.sum" not in rows:
            continue
        tot = []
        for i in cols:
            b = 0.0
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                vals, u = rows[key]
                b += float(vals[i]) * unit.get(u, 1.0)
            tot.append(b)
        return sum(tot) / len(tot), os.path.basename(path), len(tot)
    return None, None, 0


def _pin_to_gpu_numa_node(local):
    """Bind this rank (and the pinned buffers it allocates afterwards, first touch) to the CPUs of its GPU's NUMA node:
    at N = 8 eight ranks upload 25 MB per step each, and cross-socket pinned memory halves the host side of that."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        n = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def _peaks():
    p = os.path.join(paths.REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append([s.strip() for s in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for n, v in zip(names, s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def cpu_forward_timer(steps, warmup, workload="C2"):
    """Times the reference algorithm (oracle port, torch CPU fp32, all host threads) on the workload's batch."""
    sys.path.insert(0, os.path.join(paths.REPO, "tests"))
    from helpers import build_model, inputs_for
    from oracle import i2r_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = WORKLOADS[workload]
    h, w = wl.get("hw", (256, 192))
    cfg, _, sd = build_model(wl["yaml"])
    length = [wl["persons"]] * wl["images"]
    x, pm = inputs_for(length, h, w)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            i2r_oracle.forward(sd, cfg, x, pm, length)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    crops = sum(length)
    mean = sum(times) / len(times)
    return {"value": crops / mean, "unit": "crops/s", "cores": cores, "kind": "port",
            "sample": "%d forwards of the %s batch (%d crops, %dx%d), oracle port of the reference forward, "
                      "torch CPU fp32, %d threads" % (len(times), workload, crops, h, w, cores)}, mean


REF_TIMER = r'''
import json, os, sys, time
import torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "intra-and-inter-human-relation-network-for-mpee_b200"))
from oracle import ref_harness
from i2r_b200.synth import synth_inputs, synth_state_dict
yaml_rel, persons, images, h, w, steps, warmup = sys.argv[2], *map(int, sys.argv[3:9])
torch.set_num_threads(os.cpu_count() or 1)
cfg, model = ref_harness.build_reference_model(yaml_rel)
model.load_state_dict(synth_state_dict(model.state_dict(), seed=0), strict=True)
length = [persons] * images
x, pm = synth_inputs(sum(length), h, w, seed=1)
times = []
with torch.no_grad():
    for i in range(warmup + steps):
        t0 = time.perf_counter(); model(x, pm, length); dt = time.perf_counter() - t0
        if i >= warmup: times.append(dt)
print("REFTIME " + json.dumps(times))
'''


def reference_forward_timer(steps, warmup, workload="C2"):
    """Times the REAL reference (the unmodified /root/reference forward, imported through oracle/ref_harness in a
    separate process: it shares top-level module names with this repo) when the tree is present; None otherwise (the GPU
    box has no /root/reference -- the oracle port is timed there)."""
    sys.path.insert(0, paths.REPO)
    from oracle import ref_harness
    if not ref_harness.available():
        return None
    wl = WORKLOADS[workload]
    h, w = wl.get("hw", (256, 192))
    proc = subprocess.run([sys.executable, "-c", REF_TIMER, paths.REPO, wl["yaml"], str(wl["persons"]), str(wl["images"]),
                           str(h), str(w), str(steps), str(warmup)], capture_output=True, text=True)
    for line in proc.stdout.splitlines():
        if line.startswith("REFTIME "):
            times = json.loads(line[8:])
            crops = wl["persons"] * wl["images"]
            mean = sum(times) / len(times)
            cores = os.cpu_count() or 1
            return {"value": crops / mean, "unit": "crops/s", "cores": cores, "kind": "reference",
                    "sample": "%d forwards of the %s batch (%d crops, %dx%d) through the unmodified reference model "
                              "(%s, lib/models get_pose_net + forward), torch CPU fp32, %d threads" % (
                                  len(times), workload, crops, h, w, ref_harness.REF_ROOT, cores)}, mean
    return None


def run_reference(args):
    rank, world, _ = _dist_env()
    if rank != 0:
        return
    steps, warmup = min(args.steps, 10), min(args.warmup, 2)
    timed = reference_forward_timer(steps, warmup, args.workload) or cpu_forward_timer(steps, warmup, args.workload)
    base, mean = timed
    line = {"impl": "reference", "metric": "person-crops/sec", "value": base["value"], "unit": "crops/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": mean * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload]["text"] + " -- CPU forward of the reference algorithm"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--images", type=int, default=0, help="images per rank per step (default: the workload's)")
    ap.add_argument("--sharded", action="store_true",
                    help="crop-sharded forward: every rank is handed the GLOBAL batch (images x world), runs the per-crop "
                         "stages on its slice, ONE NCCL all-gather of the pooled token maps, image-complete inter-human "
                         "windows (i2r_b200/sharded.py)")
    ap.add_argument("--sustain-seconds", type=float, default=2.0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank, world, local = _dist_env()
    if args.warmup < 3:
        args.warmup = 3
    import torch.distributed as dist
    numa_cpus = _pin_to_gpu_numa_node(local) if world > 1 else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sys.path.insert(0, os.path.join(paths.REPO, "tests"))
    from helpers import build_model
    from i2r_b200.engine import HostPipeline
    from i2r_b200.synth import synth_inputs
    wl = WORKLOADS[args.workload]
    H, W = wl.get("hw", (256, 192))
    cfg, model, sd = build_model(wl["yaml"])
    model = model.cuda(dev)
    images = args.images or wl["images"]
    length = [wl["persons"]] * images              # this rank's images
    crops = sum(length)
    if args.sharded:
        from i2r_b200.sharded import ShardedForward
        call_length = length * world               # the global batch, handed to every rank
        fwd = ShardedForward(model, persons_bound=wl["persons"])
    else:
        call_length = length
        fwd = model
    call_crops = sum(call_length)

    def out_bytes(o):
        return sum(v.numel() * 4 for v in o.values()) if isinstance(o, dict) else o.numel() * 4

    # ---- inputs: NBUF distinct batches so consecutive steps do not re-read the same lines from L2
    NBUF = 8
    hx, hm = [], []
    for i in range(NBUF):
        x, pm = synth_inputs(call_crops, H, W, seed=100 + (0 if args.sharded else rank) * NBUF + i)
        hx.append(x.pin_memory())
        hm.append(pm.pin_memory())
    dx = [t.to(dev) for t in hx]
    dm = [t.to(dev) for t in hm]
    in_bytes = (hx[0].numel() + hm[0].numel()) * 4 * crops // call_crops      # what this rank uploads per step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- launches per forward (eager, counted by the Runner) and per-kernel timing of the dominant kernel
    fwd.use_cuda_graph = False
    model.use_cuda_graph = False
    fwd(dx[0], dm[0], call_length)
    torch.cuda.synchronize(dev)
    r = model._program.runner
    r.launches = 0
    r.timing = []
    for i in range(3):
        # park the GPU behind a ~10 ms spin so that the host (slower than the kernels in eager mode) runs ahead and the
        # CUDA events around each launch group bracket kernel time only, not host launch gaps
        torch.cuda._sleep(20_000_000)
        fwd(dx[i % NBUF], dm[i % NBUF], call_length)
    torch.cuda.synchronize(dev)
    launches_per_forward = r.launches // 3
    ig_ms = sum(t[0].elapsed_time(t[1]) for t in r.timing) / 3
    ig_flops = sum(t[2] for t in r.timing) / 3
    ig_launches = len(r.timing) // 3
    # dominant launch class: launches of the tcgen05 conv kernels with the same (algorithmic FLOPs, problems per group)
    # signature; the class with the largest total time is the one the roofline object describes
    classes = {}
    for a, b, f, nprob, smem in r.timing:
        c = classes.setdefault((round(f), nprob), [0.0, 0, smem])
        c[0] += a.elapsed_time(b)
        c[1] += 1
    (dom_flops, dom_nprob), (dom_ms_total, dom_n, dom_smem) = max(classes.items(), key=lambda kv: kv[1][0])
    dom_ms = dom_ms_total / dom_n
    dom_share = dom_ms_total / 3 / ig_ms if ig_ms > 0 else 0.0
    r.timing = None
    fwd.use_cuda_graph = True
    model.use_cuda_graph = True

    # ---- resident-input throughput
    for i in range(args.warmup):
        out = fwd(dx[i % NBUF], dm[i % NBUF], call_length)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        out = fwd(dx[i % NBUF], dm[i % NBUF], call_length)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # ---- the same loop sustained for >= --sustain-seconds (clocks / power settle; the 20-step region is ~40 ms)
    from i2r_b200.sharding import max_over_ranks
    ms_all = max_over_ranks([ms], device=dev)[0]      # every rank must run the same number of steps (collectives)
    sus_steps = max(args.steps, int(args.sustain_seconds * 1e3 / max(ms_all / args.steps, 1e-3)))
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    s0.record()
    for i in range(sus_steps):
        out = fwd(dx[i % NBUF], dm[i % NBUF], call_length)
    s1.record()
    barrier()
    ms_sus = s0.elapsed_time(s1)
    # ---- end-to-end: pinned host inputs -> module call -> heatmaps in pinned host memory, every step, through the
    # public pipelined entry point (engine.HostPipeline): upload of step i+1 and download of step i-1 run on their own
    # streams under the kernels of step i; the result of every step is consumed on the host (one step late)
    pipe = HostPipeline(fwd, depth=2)
    prev = None
    for i in range(3):
        t = pipe.submit(hx[i % NBUF], hm[i % NBUF], call_length)
        if prev is not None:
            prev.result()
        prev = t
    prev.result()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    prev, checksum = None, 0.0
    for i in range(args.steps):
        t = pipe.submit(hx[i % NBUF], hm[i % NBUF], call_length)
        if prev is not None:
            res = prev.result()
            checksum += float((res["multi"] if isinstance(res, dict) else res)[0, 0, 0, 0])
        prev = t
    res = prev.result()
    checksum += float((res["multi"] if isinstance(res, dict) else res)[0, 0, 0, 0])
    torch.cuda.synchronize(dev)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    clocks = sampler.finish() if rank == 0 else None

    ms, ms_e2e, ms_sus = max_over_ranks([ms, ms_e2e, ms_sus], device=dev)

    if rank == 0:
        peaks, peak_kind = _peaks()
        total = crops * world * args.steps
        value = total / (ms * 1e-3)
        e2e_value = total / (ms_e2e * 1e-3)
        sus_value = crops * world * sus_steps / (ms_sus * 1e-3)
        burst = float(peaks["bf16_tflops"])
        sustained = float(peaks.get("bf16_tflops_sustained", burst))
        achieved_all = ig_flops / (ig_ms * 1e-3) / 1e12 if ig_ms > 0 else 0.0
        achieved = dom_flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
        traffic, traffic_src, traffic_n = _ncu_traffic("conv_halo_kernel") if args.workload == "C2" else (None, None, 0)
        gf = GFLOP_PER_CROP[args.workload]
        line = {
            "metric": "person-crops/sec", "value": value, "unit": "crops/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
            "config": {"workload": wl["text"] + ("" if images == wl["images"] else " [--images %d: %d crops per GPU]" % (
                           images, crops)),
                       "partitioning": ("crop-sharded over %d ranks: global batch of %d crops handed to every rank, "
                                        "per-crop stages on the local slice, ONE all_gather_into_tensor of the pooled "
                                        "token maps (%d bytes received per rank per step), image-complete inter-human "
                                        "windows, local heads" % (world, call_crops, fwd.bytes_gathered))
                       if args.sharded else "whole images per rank, no data-path collective",
                       "l2": "inputs rotate over %d distinct batches (%.0f MB > 126 MB L2)" % (
                           NBUF, NBUF * (hx[0].numel() + hm[0].numel()) * 4 / 1e6),
                       "precision": wl["precision"],
                       "cuda_graph": True,
                       "e2e_path": "engine.HostPipeline(depth=2): pinned host inputs, staged H2D on a copy stream, D2H "
                                   "of every step's heatmaps into pinned host memory on a third stream, each result "
                                   "read on the host one step late",
                       "numa_cpus_bound": numa_cpus},
            "e2e": {"value": e2e_value, "unit": "crops/s", "h2d_bytes_per_step": in_bytes,
                    "d2h_bytes_per_step": out_bytes(out)},
            "sustained": {"value": sus_value, "unit": "crops/s", "seconds": ms_sus * 1e-3, "steps": sus_steps},
            "gpu_launches": launches_per_forward * args.steps,
            "model_tflops": value * gf / 1e3,
            "model_frac_of_peak": value * gf / 1e3 / (burst * world),
            "model_frac_of_sustained_peak": sus_value * gf / 1e3 / (sustained * world),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": burst, "unit": "TFLOP/s",
                         "frac": achieved / burst, "traffic": traffic,
                         "traffic_source": None if traffic is None else "%s (%d launches)" % (traffic_src, traffic_n),
                         "frac_of_sustained_peak": achieved / sustained,
                         "kernel": "conv_halo_kernel, dominant launch class (%d problems per launch, %.2f algorithmic "
                                   "GFLOP per launch = sum 2*M*Cout*Cin*taps; %d launches per forward, %.0f%% of the "
                                   "conv-kernel time); average launch duration %.1f us from CUDA events around every "
                                   "launch of an eager forward queued behind a spin kernel (host gaps excluded)" % (
                                       dom_nprob, dom_flops / 1e9, dom_n // 3, 100 * dom_share, dom_ms * 1e3),
                         "all_conv_launches": {"achieved": achieved_all, "frac": achieved_all / burst,
                                               "launch_groups_per_forward": ig_launches},
                         "smem_port": _smem_port(dom_smem, dom_ms, clocks, dev),
                         "peak_kind": "bf16_tflops (burst: per-launch event timing), %s; sustained %.1f kept in "
                                      "frac_of_sustained_peak" % (peak_kind, sustained)},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            base, _ = cpu_forward_timer(3, 1, args.workload)
            line["cpu_baseline"] = base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
