#!/usr/bin/env python
"""bench.py - the head hot path on N B200s (one process per GPU; torchrun for N > 1).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle)

A "step" is one pass of the hot path over one synthetic batch (BASELINE.json configs[1]):
B=16 images per GPU at 512^2, D=512, C=150, bf16 GEMM operands / fp32 accumulate:
cosine logits (K0+K1) -> fused bilinear-upsample + softmax-CE fwd/bwd (K2) -> logits backward
(K1b) -> argmax / confusion matrix / mIoU (K3); with N > 1 the valid-pixel count, the head-gradient
bucket and the int64 confusion matrix are all-reduced over NCCL every step (weak scaling).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ADE20K-shape 512^2 images/sec (logits+loss+mIoU)"
UNIT = "images/s"
GEOM = {"A": (32, 32, 512), "B": (128, 128, 512), "5": (64, 64, 1024)}   # (h, w, H = W): SURVEY 8 G-A aux head x16,
#                                   G-B main head x4; "5" = BASELINE config 5 (1024^2, use --classes 847 --batch 8)
L2_BYTES = 126 * 1024 * 1024


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=None)
    p.add_argument("--warmup", type=int, default=None)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--geometry", default="A", choices=list(GEOM))
    p.add_argument("--batch", type=int, default=16, help="images per GPU")
    p.add_argument("--classes", type=int, default=150)
    p.add_argument("--no-backward", action="store_true")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--kernel-times", action="store_true", help="extra pass: per-section CUDA-event times (not the timed run)")
    return p.parse_args()


def workload_name(a, h, w):
    H = GEOM[a.geometry][2]
    return (f"cfg{'5' if a.geometry == '5' else '2'}: text_patch logits + fused upsample/softmax-CE fwd/bwd + argmax/confmat mIoU, "
            f"B={a.batch}/GPU, {h}x{w}->{H}x{H} (x{H // h}), D=512, C={a.classes}, ignore_index=0")


# --------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Polls NVML (SM clock, throttle reasons) while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:  # noqa: BLE001
            return local_rank
    return local_rank


# --------------------------------------------------------------------------------------------
def cpu_reference_time(B, h, w, C, steps, warmup, threads=None):
    """Times oracle.head_step (the reference's lines on CPU, fp32) on a B-image sample."""
    import torch
    from lc2is_b200 import synthetic
    from oracle import head_oracle as O
    if threads:
        torch.set_num_threads(threads)
    v = synthetic.make_patch_embeddings(B, h * w, 512, dtype=torch.float32)
    t = synthetic.make_prototypes(C, 512)
    labels = synthetic.make_labels(B, 512, 512, C)
    for _ in range(warmup):
        O.head_step(v, t, labels, ignore_index=0, n_cls=C, hw_shape=(h, w))
    t0 = time.perf_counter()
    for _ in range(steps):
        O.head_step(v, t, labels, ignore_index=0, n_cls=C, hw_shape=(h, w))
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dt, torch.get_num_threads()


def run_reference(a):
    """--impl reference: the reference's own CPU implementation of the path (the oracle: reference lines
    restated with live torch ops; the reference is pure Python and cannot run as shipped - SURVEY 0) on
    all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    h, w, HW_ = GEOM[a.geometry]
    steps = a.steps if a.steps is not None else 5
    warmup = a.warmup if a.warmup is not None else 1
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = 2
    t1, _ = cpu_reference_time(Bs, h, w, a.classes, 1, 0, cores)          # probe
    if t1 * (steps + warmup) > 150 and Bs > 1:
        Bs = 1
    dt, thr = cpu_reference_time(Bs, h, w, a.classes, steps, warmup, cores)
    val = Bs / dt
    sample = f"{Bs} images per step of the same workload (oracle.head_step, fp32, torch CPU)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a, h, w), "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": thr, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist
    from lc2is_b200 import _lib, dp, synthetic
    from lc2is_b200.step import HeadStep, HostStep

    rank, world, local_rank = dp.init_distributed()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: lc2is_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    steps = a.steps if a.steps is not None else 200
    warmup = a.warmup if a.warmup is not None else 20
    warmup = max(warmup, 3)
    h, w, HW_ = GEOM[a.geometry]
    B, C, D, H, W = a.batch, a.classes, 512, HW_, HW_
    backward = not a.no_backward

    # ---- inputs: rotate over enough distinct sets that a step's inputs never sit in the 126 MB L2
    set_bytes = B * h * w * D * 2 + B * H * W * 8
    nset = max(2, -(-2 * L2_BYTES // set_bytes))
    t_host = synthetic.make_prototypes(C, D)
    host_sets = []
    for i in range(nset):
        hv = synthetic.make_patch_embeddings(B, h * w, D, seed=synthetic.SEED + 97 * (rank * nset + i)).pin_memory()
        hl = synthetic.make_labels(B, H, W, C, seed=synthetic.SEED + 97 * (rank * nset + i), ignore_frac=0.1).pin_memory()
        host_sets.append((hv, hl))
    dev_sets = [(hv.to(dev), hl.to(dev)) for hv, hl in host_sets]
    t_dev = t_host.to(dev)
    t_pin = t_host.pin_memory()

    step = HeadStep(B, h, w, H, W, C, D, ignore_index=0, backward=backward, device=dev, distributed=world > 1)
    k2_pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- device-resident timing -----------------------------------------------------------------------
    for i in range(warmup):
        step(dev_sets[i % nset][0], t_dev, dev_sets[i % nset][1])
    step.finish()
    torch.cuda.synchronize()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    barrier()
    torch.cuda.synchronize()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step.reset_metrics()
    sampler.start()
    t_host0 = time.perf_counter()
    ev0.record()
    for i in range(steps):
        step.k2_events = k2_pairs[i]
        v_i, l_i = dev_sets[(warmup + i) % nset]
        step(v_i, t_dev, l_i)
    cm_total = step.global_confmat()                         # ONE int64 all-reduce for the whole pass (N > 1)
    ev1.record()
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / steps     # host time to enqueue one step
    torch.cuda.synchronize()
    sampler.stop_flag = True
    barrier()
    step.k2_events = None
    launches = _lib.launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    k2_ms = statistics.mean(s.elapsed_time(e) for s, e in k2_pairs)
    tm = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms_total = float(tm)
    ms_per_step = ms_total / steps
    value = world * B * steps / (ms_total * 1e-3)
    loss_val = float(step.loss)
    miou = None
    try:
        from lc2is_b200 import metrics
        miou = float(metrics.miou_from_confmat(cm_total, 0))
    except Exception:  # noqa: BLE001
        pass
    sampler.join(timeout=1)
    clocks = sampler.result()

    section_us = None
    if a.kernel_times:
        step.timers = {}
        for i in range(20):
            step(dev_sets[i % nset][0], t_dev, dev_sets[i % nset][1])
        step.finish()
        torch.cuda.synchronize()
        section_us = {k: round(statistics.mean(s.elapsed_time(e) for s, e in v) * 1e3, 1) for k, v in step.timers.items()}
        step.timers = None

    # ---- end to end: host (pinned) buffers through the C-ABI host entry, H2D/D2H inside the timed region
    e2e = None
    if not a.no_e2e:
        hstep = HostStep(B, h, w, H, W, C, D, ignore_index=0, backward=backward, device=dev, depth=2)
        cm_dev = torch.zeros(C, C, dtype=torch.int64, device=dev)
        e_steps = max(3, min(steps, 200))

        cm_host = torch.zeros(C, C, dtype=torch.int64)

        def finish(out):
            cm_host.add_(out[2])                            # the pass's confusion matrix accumulates on the host

        def reduce_confmat():
            if world > 1:                                   # DP: ONE int64 all-reduce for the whole pass
                cm_dev.copy_(cm_host)
                dist.all_reduce(cm_dev)
                torch.cuda.synchronize()

        def run(n, first):
            """n steps through submit / wait with two steps in flight: every step copies its own inputs from
            pinned host memory and reads its own results back; the copy of step i+1 overlaps the kernels of
            step i (a prefetching loader).  Returns the last results."""
            hv, hl = host_sets[first % nset]
            hstep.submit(hv, t_pin, hl)
            out = None
            for i in range(1, n):
                hv, hl = host_sets[(first + i) % nset]
                hstep.submit(hv, t_pin, hl)
                if i + 1 < n:                               # loader-style prefetch: the library's host threads pack the
                    hstep.prefetch(host_sets[(first + i + 1) % nset][1])   # next batch's labels while this thread waits
                out = hstep.wait()
                finish(out)
            out = hstep.wait()
            finish(out)
            reduce_confmat()
            return out

        def timed(fn):
            barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            barrier()
            ms = max(e0.elapsed_time(e1), wall * 1e3)        # the calls block the host: take the larger clock
            tm = torch.tensor([ms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            return float(tm)

        run(4, 0)
        ms_pipe = timed(lambda: run(e_steps, 4))
        assert int(hstep.out_n_valid) > 0

        def blocking():
            for i in range(e_steps):
                hv, hl = host_sets[(4 + i) % nset]
                hstep(hv, t_pin, hl)
                finish((hstep.out_loss, hstep.out_n_valid, hstep.out_confmat))
            reduce_confmat()
        hstep(host_sets[0][0], t_pin, host_sets[0][1])
        ms_block = timed(blocking)
        e2e = {"value": world * B * e_steps / (ms_pipe * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": hstep.h2d_bytes, "d2h_bytes_per_step": hstep.d2h_bytes,
               "steps": e_steps, "ms_per_step": ms_pipe / e_steps,
               "api": "lc2is_head_step_host_submit / _wait, 2 steps in flight, labels of the next batch packed in the "
                      "background (lc2is_pack_labels_host_begin/_end); pinned host buffers in: bf16 V, fp32 T, int64 labels "
                      "narrowed to 1 byte (C <= 254) by the library's host threads and widened on the device; loss/n_valid/confmat out",
               "label_route": {"host_label_bytes": hstep.label_bytes, "n_raw_int64_images": hstep.n_raw if hstep.host_pack else B,
                               "calibration": hstep.calibration},
               "blocking_call": {"api": "lc2is_head_step_host", "ms_per_step": ms_block / e_steps,
                                 "value": world * B * e_steps / (ms_block * 1e-3)}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (K2: fused upsample + CE fwd/bwd) ------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    split, fused = step.split, getattr(step, "fused", False)
    # Dominant launch: on the x16 geometry K2 (fused upsample + softmax-CE fwd/bwd) and K3 (argmax + confusion matrix)
    # are ONE warp-specialised kernel.  Algorithmic bytes of that launch: read low + accumulate grad_low (fp32) + the
    # packed uint16 label map (read by the CE warps and by the argmax warps) + the loss scalar.
    if fused:
        k2_bytes = 2 * B * C * h * w * 4 + 2 * B * H * W * 2 + 8
        k2_name = "k23_fused_kernel<16>"
        k2_desc = "lc2is_ce_argmax_fused_packed: bilinear upsample + softmax-CE fwd/bwd + argmax + confusion matrix"
    else:
        k2_bytes = 2 * B * C * h * w * 4 + B * H * W * (2 if split else 8) + 8
        k2_name = ("k2_strip_kernel<16, true>" if H // h == 16 else "k2_strip_kernel<8, true>") if split else \
            "k2_fast_kernel / k2_generic_kernel"
        k2_desc = "lc2is_upsample_ce_packed: fused bilinear upsample + softmax-CE fwd+bwd"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and a.geometry == "A" and B == 16 and C == 150:
        traffic = json.load(open(tpath)).get(k2_name)                 # from one `ncu --set full` capture
    achieved = k2_bytes / (k2_ms * 1e-3) / 1e9
    roofline = {"kernel": f"{k2_name} ({k2_desc})",
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes": k2_bytes, "kernel_us": k2_ms * 1e3,
                "note": "instruction-bound, not HBM-bound (the upsampled [B,C,H,W] tensor never exists): see "
                        f"`compute` and DESIGN.md; {B * C * H * W / (k2_ms * 1e-3) / 1e12:.3f} T upsampled elements/s"}
    # The bounds that do bind.  CE part: 4 packed fp32x2 instructions per pixel pair and class (pass A: add + mul,
    # pass B: two fma) at the measured packed issue rate of 2.0 warp-inst/clk/SM (tools/micro/k2loops.cu).
    # Argmax part (fused kernel only): one FFMA2 and one FMNMX3 per pixel pair and class at 4 issue slots/clk/SM.
    sm_hz = (clocks.get("sm_mhz") or 1965) * 1e6
    ce_inst = B * C * H * W / 2 * 4 / 32
    ce_us = ce_inst / (2.0 * 148 * sm_hz) * 1e6
    am_inst = B * C * H * W / 2 * 2 / 32 if fused else 0.0
    am_us = am_inst / (4.0 * 148 * sm_hz) * 1e6
    roofline["compute"] = {"ce_packed_fp32_warp_inst": ce_inst, "ce_us_at_2_per_clk_per_sm": ce_us,
                           "argmax_warp_inst": am_inst, "argmax_us_at_4_per_clk_per_sm": am_us,
                           "frac": (ce_us + am_us) / (k2_ms * 1e3)}

    cpu_baseline = None
    if not a.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        dt, thr = cpu_reference_time(2, h, w, C, 2, 1, cores)
        cpu_baseline = {"value": 2 / dt, "unit": UNIT, "cores": thr, "kind": "port",
                        "sample": "2 images per step of the same workload, 2 timed steps after 1 warm-up "
                                  "(oracle.head_step: reference lines on torch CPU, fp32)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(a, h, w), "global_batch": world * B, "parallelism": f"dp{world}",
                   "backward": backward,
                   "l2": f"inputs rotate over {nset} distinct sets ({nset * set_bytes / 2**20:.0f} MiB) > 126 MiB L2; no flush"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        "cpu_baseline": cpu_baseline, "section_us": section_us, "host_enqueue_ms_per_step": host_enqueue_ms,
        "check": {"loss": loss_val, "mIoU": miou, "n_valid": int(step.n_valid)},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
