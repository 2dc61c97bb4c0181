#!/usr/bin/env python
"""Benchmark of the DCT-domain ViT hot path on B200 (contract: see the task's bench section).

    python bench.py --gpus 1 --steps 20 --warmup 5            # our arm
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1   # reference CPU path (oracle port)
    torchrun ... bench.py --gpus N ...                        # one rank per GPU, weak scaling

A "step" is one pass of the hot path over one batch of 256 synthetic 512x512 4:2:0 images per
GPU whose Huffman-decoded int16 coefficients are the input: fused DCT-aug+embed kernel (K0)
-> ViT forward/backward/optimizer when `--stage train` (default once the model is built).
`value` times the step with inputs resident in HBM; `e2e` includes the pinned-host -> device
copy of the coefficients + plans and a device -> host read of the step result.
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

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "images/sec (512x512 JPEG, ViT-S DCT) at 1/2/4/8 B200; DCT-aug+embed HBM GB/s"
UNIT = "images/s"
N_POOL = 4          # distinct input batches cycled through (4 x 201 MB >> 126 MB L2)
K0_NCU_FILES = {"train": "profiles/r02_k0_v2_train_ncu.txt", "eval": "profiles/r02_k0_v2_eval_ncu.txt"}
VIT_TRAIN_GFLOP_PER_IMAGE = {"vits": 27.28, "vitti": 7.40}      # dense GEMM + attention MACs x 2, fwd + bwd


def k0_ncu_traffic(kind: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE K0 launch (batch 256) from the committed `ncu --set full` summary of the
    current kernel (tools/ncu_summary.py writes the `traffic(dram read+write) = N bytes` line) -- read from the file, not a constant."""
    path = os.path.join(ROOT, K0_NCU_FILES[kind])
    try:
        with open(path) as f:
            for ln in f:
                if ln.strip().startswith("traffic(dram read+write)"):
                    return int(float(ln.split("=")[1].split()[0])), K0_NCU_FILES[kind]
    except OSError:
        pass
    return None, K0_NCU_FILES[kind] + " (missing)"


def ncu_tensor_pipe():
    """sm__pipe_tensor_cycles_active (% of peak sustained active) of the tcgen05 kernels, read from the committed `ncu --set full`
    summaries (captured under the profiler, never timed there): {label: percent}."""
    out = {}
    for rel in ("profiles/r01_attn_final_ncu.txt", "profiles/r01_gemm_final_ncu.txt"):
        label = None
        try:
            with open(os.path.join(ROOT, rel)) as f:
                for ln in f:
                    if ln.startswith("launch "):
                        label = ln.split(":", 1)[1].strip()
                    elif "sm__pipe_tensor_cycles_active" in ln and label is not None:
                        out[label] = round(float(ln.split("=")[1].split()[0]), 1)
        except OSError:
            out[rel] = "missing"
    return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# Workload construction
# ----------------------------------------------------------------------------------------------
def build_inputs(batch: int, rank: int, from_jpeg: int = 8):
    """N_POOL synthetic coefficient batches in pinned host memory.  The first `from_jpeg` images of
    batch 0 are real PIL-encoded JPEGs run through our Huffman decoder; the rest use the
    JPEG-statistics generator (encoding 1024 JPEGs with PIL would take minutes of setup)."""
    from rgb_no_more_b200 import dct_manip as dm, synth
    pool, flags = [], []
    for k in range(N_POOL):
        y, c, q = synth.synth_coefficients(batch, 64, 64, seed=synth.SEED + 1000 * rank + k)
        fl = synth.dequant_clamp_flags(y, c, q)          # the decoder reports this per file; generated data: computed
        y, c, q = torch.from_numpy(y), torch.from_numpy(c), torch.from_numpy(q)
        if k == 0 and from_jpeg:
            jy, jc, jq, jfl = dm.decode_batch(synth.synth_jpeg_set(from_jpeg), 64, 64, nthreads=0)
            y[:from_jpeg], c[:from_jpeg], q[:from_jpeg] = jy, jc, jq
            fl[:from_jpeg] = np.asarray(jfl)
        pool.append(tuple(t.pin_memory() for t in (y, c, q)))
        flags.append(fl)
    return pool, flags


def host_decode_rate(n_files: int = 64, repeats: int = 3):
    """Huffman-decode leg (row a1, stays on the host by north_star): rgbnm_jpeg_decode_batch on all host cores over
    `n_files` synthetic 512x512 Q75 4:2:0 JPEGs -- the rate at which this box can turn JPEG bytes into the coefficient
    batches the GPU path consumes."""
    from rgb_no_more_b200 import dct_manip as dm, synth
    jpegs = synth.synth_jpeg_set(n_files)
    y, c, q, _ = dm.decode_batch(jpegs, 64, 64, nthreads=0)        # staging buffers are reused, like a feeder ring would
    t0 = time.perf_counter()
    for _ in range(repeats):
        dm.decode_batch(jpegs, 64, 64, nthreads=0, out=(y, c, q))
    dt = (time.perf_counter() - t0) / repeats
    return {"value": n_files / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "jpeg_bytes_mean": int(np.mean([len(j) for j in jpegs])),
            "what": f"host Huffman decode of {n_files} synthetic 512x512 JPEGs into pinned-layout int16 planes, all cores; the "
                    "from-JPEG-bytes rate of one rank is min(this / ranks per host, e2e)"}


def algorithmic_bytes(plans, out_bytes: int) -> int:
    """int16 coefficients inside the crop window + 384 B tables + plan, + output."""
    total = 0
    for pl in plans:
        s = pl.crop_size
        total += s * s * 128 + 2 * (s // 2) * (s // 2) * 128 + 384 + 112 + 196 * 384 * out_bytes
    return total


def run_ours(args):
    import torch.distributed as dist
    from rgb_no_more_b200 import plan as P, transforms as TF
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    out_dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    tf = TF.FusedDCT(dev, "train", P.AUGLIST_VITS, 2, 9, out_dtype=out_dtype)
    tf_eval = TF.FusedDCT(dev, "test", out_dtype=out_dtype)
    host_pool, clamp_flags = build_inputs(B, rank)
    torch.manual_seed(11997733 + rank)
    plan_pool = [tf.sample_plans(B) for _ in range(N_POOL)]
    packed_pool = [torch.from_numpy(P.pack_plans(p, fl).view(np.uint8).reshape(B, -1).copy()).pin_memory()
                   for p, fl in zip(plan_pool, clamp_flags)]
    dev_pool = [tuple(t.to(dev) for t in hp) for hp in host_pool]
    dev_plans = [p.to(dev) for p in packed_pool]
    stage = None
    if args.stage == "train":
        from rgb_no_more_b200 import train_step as TS
        stage = TS.TrainStage(dev, arch=args.arch, batch=B, dtype=args.dtype, world=world, use_graph=not args.no_graph, rank=rank)
    # K0 writes the operand of the patch projection straight into the train stage's static input buffer
    out_buf = stage.x_static if (stage is not None and out_dtype == torch.bfloat16) else torch.empty((B, 196, 384), dtype=out_dtype, device=dev)
    labels_pool = [torch.randint(0, 1000, (B,), device=dev) for _ in range(N_POOL)]

    k0_ms = []

    def step_device(i, timed=False):
        k = i % N_POOL
        y, c, q = dev_pool[k]
        x = tf.run(y, c, q, None, plans_dev=dev_plans[k], out=out_buf, timing=k0_ms if timed else None)
        if stage is not None:
            return stage.step(x, labels_pool[k])
        return x

    # ---- end-to-end arm: host buffers in, host scalar out, every step ---------------------------------------
    # Two device staging sets; the pinned-host -> device copy of step i+1 runs on a copy stream while step i computes
    # (the copy engine and the SMs overlap), each timed region still contains exactly one copy per step: the first
    # copy of a region is exposed, the last step of a region does not prefetch.
    h2d_stream = torch.cuda.Stream(device=dev)
    stage_bufs = [tuple(torch.empty_like(t, device=dev) for t in host_pool[0]) for _ in range(2)]
    stage_plans = [torch.empty_like(packed_pool[0], device=dev) for _ in range(2)]
    result_host = torch.empty((1,), dtype=torch.float32).pin_memory()
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    copied = set()
    region_last = {args.warmup - 1, args.warmup + args.steps - 1}

    def issue_copy(i):
        k, sidx = i % N_POOL, i % 2
        with torch.cuda.stream(h2d_stream):
            h2d_stream.wait_event(ev_free[sidx])              # the step that last read this staging set has finished
            for dst, src in zip(stage_bufs[sidx], host_pool[k]):
                dst.copy_(src, non_blocking=True)
            stage_plans[sidx].copy_(packed_pool[k], non_blocking=True)
            ev_ready[sidx].record(h2d_stream)
        copied.add(i)

    def step_e2e(i):
        k, sidx = i % N_POOL, i % 2
        if i not in copied:
            issue_copy(i)
        if i not in region_last:
            issue_copy(i + 1)
        cur = torch.cuda.current_stream()
        cur.wait_event(ev_ready[sidx])
        x = tf.run(*stage_bufs[sidx], None, plans_dev=stage_plans[sidx], out=out_buf)
        if stage is not None:
            res = stage.step(x, labels_pool[k])
        else:
            res = x[:, 0, :8].float().sum().reshape(1)     # a scalar that depends on the step's output
        ev_free[sidx].record(cur)
        result_host.copy_(res.reshape(-1)[:1].float(), non_blocking=True)
        copied.discard(i)

    def timed_loop(fn, steps, warmup, **kw):
        for i in range(warmup):
            fn(i, **kw) if kw else fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i, **kw) if kw else fn(warmup + i)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed_loop(step_device, args.steps, args.warmup, timed=True)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed_loop(step_e2e, args.steps, args.warmup)

    # ---- from JPEG bytes: host Huffman decode -> pinned ring -> copy stream -> K0 -> step, at ANY world size -------------------
    # Every rank decodes its own batches with cpu_count // world host threads (the reference gives each rank num_cpus // world_size
    # DataLoader workers, pipeline_utils.py:125) and draws fresh augmentation plans for every batch INSIDE the timed region
    # (plan.BatchedSampler: same distributions as the per-image sampler, one vectorised draw per batch).
    e2e_jpeg = None
    if not args.no_cpu and stage is not None:
        from rgb_no_more_b200 import feeder as FD, synth
        n_threads = max(1, (os.cpu_count() or 1) // world)
        if n_threads >= 8:
            n_threads -= 1                            # leave a core to the thread that feeds the GPU
        jpegs = synth.synth_jpeg_set(min(B, 64))
        jpegs = (jpegs * (B // len(jpegs) + 1))[:B]
        fd = FD.JpegFeeder(dev, B, 64, 64, slots=3, nthreads=n_threads)
        sampler_b = P.BatchedSampler(64, 64, P.AUGLIST_VITS, 2, 9, tf.bank)
        plan_host = [torch.empty((B, P.PLAN_DTYPE.itemsize), dtype=torch.uint8).pin_memory() for _ in range(3)]
        plan_dev = [torch.empty((B, P.PLAN_DTYPE.itemsize), dtype=torch.uint8, device=dev) for _ in range(3)]
        gen = torch.Generator().manual_seed(11997733 + rank)
        plan_ev = [None, None, None]

        drawn = []

        def step_jpeg(i):
            while len(fd.pending) < 2:
                # plan first, decode second: the decoder stops after the last block row a crop window needs
                pk = sampler_b.sample(B, generator=gen)
                drawn.append(pk)
                fd.submit(jpegs, last_rows=pk["crop_i"].astype(np.int32) + pk["crop_size"].astype(np.int32) - 1)   # overlaps this step
            yj, cj, qj, flags, slot = fd.get()
            k = i % 3
            packed = drawn.pop(0)
            packed["clamp_in"] = flags
            if plan_ev[k] is not None:
                plan_ev[k].synchronize()              # the copy that last read this pinned buffer (3 steps ago) has completed
            plan_host[k].copy_(torch.from_numpy(packed.view(np.uint8).reshape(B, -1)))
            plan_dev[k].copy_(plan_host[k], non_blocking=True)
            if plan_ev[k] is None:
                plan_ev[k] = torch.cuda.Event()
            plan_ev[k].record()
            x = tf.run(yj, cj, qj, None, plans_dev=plan_dev[k], out=out_buf)
            fd.release(slot)
            res = stage.step(x, labels_pool[i % N_POOL])
            result_host.copy_(res.reshape(-1)[:1].float(), non_blocking=True)

        n_j = max(args.steps // 2, 5)
        ms_jpeg = timed_loop(step_jpeg, n_j, 3)
        while fd.pending:
            _, _, _, _, slot = fd.get()
            fd.release(slot)
        torch.cuda.synchronize()
        fd.close()
        e2e_jpeg = {"value": B * world / (ms_jpeg / n_j * 1e-3), "unit": UNIT, "jpeg_bytes_per_step": int(sum(len(j) for j in jpegs)),
                    "h2d_bytes_per_step": int(fd.h2d_bytes) + B * P.PLAN_DTYPE.itemsize, "host_cores": os.cpu_count() or 1,
                    "decode_threads_per_rank": n_threads, "n_gpus": world,
                    "what": "whole job fed from JPEG byte strings: every rank runs rgbnm_jpeg_decode_batch on cpu_count // world host "
                            "threads into a pinned ring (plan first: the scan is abandoned after the crop window's last block row), copy "
                            "stream, fresh plans per batch (BatchedSampler, inside the timed region), K0, train step; decode of the next "
                            "batches overlaps the GPU"}

    # K0 kernel(s) alone: average over the timed region (events on the launching stream)
    k0_avg_ms = float(np.mean([a.elapsed_time(c) for a, _, c in k0_ms[-args.steps:]]))         # statistics pre-pass + fused kernel
    k0_fused_ms = float(np.mean([b.elapsed_time(c) for _, b, c in k0_ms[-args.steps:]]))      # the fused kernel's own launches
    out_bytes = 2 if out_dtype == torch.bfloat16 else 4
    alg = float(np.mean([algorithmic_bytes(p, out_bytes) for p in plan_pool]))
    pk, pk_src = peaks()
    # canonical eval geometry side measurement
    eplans = [torch.from_numpy(P.pack_plans(tf_eval.sample_plans(B), fl).view(np.uint8).reshape(B, -1).copy()).to(dev)
              for fl in clamp_flags]

    # N_POOL launches over the N_POOL distinct input batches captured in one CUDA graph: the replay is not limited by
    # the host's launch rate (two launches + two small allocations per call from Python cost about as much as the kernel)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for k in range(N_POOL):
            tf_eval.run(*dev_pool[k], None, plans_dev=eplans[k], out=out_buf)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g_eval = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_eval):
        for k in range(N_POOL):
            tf_eval.run(*dev_pool[k], None, plans_dev=eplans[k], out=out_buf)

    def step_eval(i):
        g_eval.replay()
    n_eval = max(args.steps, 20)
    ms_eval = timed_loop(step_eval, n_eval, max(args.warmup, 3)) / (n_eval * N_POOL)
    alg_eval = algorithmic_bytes(tf_eval.sample_plans(B), out_bytes)

    # eval-only forward (eval.py:36-38): K0 in the eval geometry -> ViT forward, no gradients
    eval_fwd_ips = None
    if stage is not None:
        x_eval = torch.empty((B, 196, 384), dtype=torch.bfloat16, device=dev)
        with torch.no_grad():
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                tf_eval.run(*dev_pool[0], None, plans_dev=eplans[0], out=x_eval)
                stage.eng.forward(x_eval, save=False)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g_fwd = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_fwd):
                for k in range(N_POOL):
                    tf_eval.run(*dev_pool[k], None, plans_dev=eplans[k], out=x_eval)
                    stage.eng.forward(x_eval, save=False)
        n_fwd = max(args.steps // 2, 5)
        ms_fwd = timed_loop(lambda i: g_fwd.replay(), n_fwd, 3) / (n_fwd * N_POOL)
        eval_fwd_ips = B * world / (ms_fwd * 1e-3)

    # SwinV2-T DCT eval forward (BASELINE config 5 shape, per GPU): K0 in the Swin layout
    # (Resize_DCT(32) geometry) -> SwinTransformerV2 forward, same resident coefficient batches, CUDA-graph replay
    swin_fwd = None
    if world == 1 and args.stage == "train" and out_dtype == torch.bfloat16 and not args.no_swin:
        try:
            swin_fwd = swin_eval_forward(args, dev, dev_pool, clamp_flags, timed_loop, host_pool)
        except Exception as ex:  # noqa: BLE001 -- a side measurement must never cost the headline line
            swin_fwd = {"error": f"{type(ex).__name__}: {ex}"[:300]}

    # SwinV2-T DCT TRAINING (BASELINE config 5: window 8, bf16, data parallel), at every world size: K0 in the Swin layout with the
    # RandAugment mix -> SwinTransformerV2 forward / backward (rgb_no_more_b200.swin_train) -> DDP all-reduce -> AdamW
    swin_train = None
    if args.stage == "train" and out_dtype == torch.bfloat16 and not args.no_swin:
        try:
            swin_train = swin_train_step(args, dev, dev_pool, timed_loop, rank, world, local)
        except Exception as ex:  # noqa: BLE001
            swin_train = {"error": f"{type(ex).__name__}: {ex}"[:300]}
    # BASELINE configs 1 and 2 (ViT-Ti) as side measurements on one GPU
    vitti = None
    if world == 1 and args.stage == "train" and args.arch == "vits" and out_dtype == torch.bfloat16 and not args.no_cpu:
        try:
            vitti = vitti_configs(args, dev, dev_pool, dev_plans, tf, tf_eval, labels_pool, timed_loop)
        except Exception as ex:  # noqa: BLE001
            vitti = {"error": f"{type(ex).__name__}: {ex}"[:300]}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    ms_step = ms_total / args.steps
    value = B * world / (ms_step * 1e-3)
    e2e_val = B * world / (ms_e2e / args.steps * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in host_pool[0]) + packed_pool[0].numel()
    launches = 2 + (stage.launches_per_step if stage is not None else 0)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": (f"{args.arch} DCT train step (K0 fused DCT-aug+embed -> ViT fwd/bwd/AdamW)" if stage is not None
                                else "K0 fused DCT-aug+embed only (ViT stage not built yet)") +
                               f", batch {B}/GPU, 512x512 4:2:0 coefficients, RandAugment num_ops=2 magnitude=9",
                   "l2_policy": f"{N_POOL} distinct input batches x 201 MB cycled (> 126 MB L2)", "batch_per_gpu": B},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4},
        "gpu_launches": launches * args.steps,
        "clocks": clocks,
        # the dominant kernel of the data path = k0_vit2_kernel, timed by events around its own launches inside the timed steps
        # (training mix of crop sizes and ops); `with_prepass` adds the DC-statistics launch in front of it
        "roofline": {"bound": "hbm", "kernel": "k0_vit2_kernel (train mix), launches inside the timed steps",
                     "achieved": alg / (k0_fused_ms * 1e-3) / 1e9,
                     "peak": pk["hbm_gbs"], "peak_source": pk_src, "unit": "GB/s",
                     "frac": alg / (k0_fused_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                     # dram__bytes_read.sum + dram__bytes_write.sum of one train-mix launch (ncu --set full, batch 256)
                     "traffic": k0_ncu_traffic("train")[0] if B == 256 else None, "traffic_source": k0_ncu_traffic("train")[1],
                     "algorithmic_bytes_per_launch": alg, "kernel_ms": k0_fused_ms,
                     "frac_of_nominal_8000_gbs": alg / (k0_fused_ms * 1e-3) / 1e9 / 8000.0,          # north_star quotes the 8 TB/s figure
                     "with_prepass": {"kernel": "k0_dcstats + k0_vit2_kernel", "kernel_ms": k0_avg_ms,
                                      "achieved": alg / (k0_avg_ms * 1e-3) / 1e9, "frac": alg / (k0_avg_ms * 1e-3) / 1e9 / pk["hbm_gbs"]}},
        "roofline_eval_geometry": {"bound": "hbm", "achieved": alg_eval / (ms_eval * 1e-3) / 1e9, "peak": pk["hbm_gbs"],
                                   "unit": "GB/s", "frac": alg_eval / (ms_eval * 1e-3) / 1e9 / pk["hbm_gbs"],
                                   "traffic": k0_ncu_traffic("eval")[0] if B == 256 else None,
                                   "traffic_source": k0_ncu_traffic("eval")[1],
                                   "algorithmic_bytes_per_launch": alg_eval, "kernel_ms": ms_eval,
                                   "frac_of_nominal_8000_gbs": alg_eval / (ms_eval * 1e-3) / 1e9 / 8000.0,
                                   "images_per_s": B / (ms_eval * 1e-3)},
    }
    if e2e_jpeg is not None:
        line["e2e_from_jpeg"] = e2e_jpeg
    if eval_fwd_ips is not None:
        line["eval_forward"] = {"value": eval_fwd_ips, "unit": UNIT, "what": "K0 (eval geometry) + ViT forward, no gradients, coefficients "
                                "resident in HBM, CUDA-graph replay"}
    if swin_fwd is not None:
        line["swin_eval_forward"] = swin_fwd
    if swin_train is not None:
        line["swin_train"] = swin_train
    if vitti is not None:
        line["vitti_configs"] = vitti
    if stage is not None and args.arch in VIT_TRAIN_GFLOP_PER_IMAGE:
        # the ViT part of the step against the measured cuBLAS bf16 rate (sustained figure: timed inside a long step)
        tf_s = VIT_TRAIN_GFLOP_PER_IMAGE[args.arch] * B / ((ms_step - k0_avg_ms) * 1e-3) / 1e3
        peak_tf = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        line["roofline_vit_step"] = {"bound": "tensor", "achieved": tf_s, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf_s / peak_tf,
                                     "note": "algorithmic GEMM + attention FLOPs of forward + backward over the whole ViT part of the "
                                             "step (LayerNorm, optimiser, bias sums included in the time, not in the FLOPs)"}
    if stage is not None:
        line["ncu_tensor_pipe_active_pct"] = {"source": "profiles/r01_attn_final_ncu.txt, profiles/r01_gemm_final_ncu.txt (ncu --set full; "
                                              "the ViT kernels are unchanged since)", **ncu_tensor_pipe()}
    if world == 1 and not args.no_cpu:
        line["host_decode"] = host_decode_rate()
        line["cpu_baseline"] = cpu_baseline(args)
        if stage is not None:
            line["torch_b200_baseline"] = torch_b200_baseline(args)
    print(json.dumps(line))


def swin_train_step(args, dev, dev_pool, timed_loop, rank, world, local):
    """BASELINE config 5 (SwinV2-T DCT, window 8, bf16, data parallel): batch 64 per GPU (reference: 512 over 8 GPUs, configs.py:138),
    drop_path 0.2, mixup 0.2, fresh plans per step from the batched sampler; train_step.TrainStage(arch='swinv2t'): K0 in the Swin
    layout -> mixup -> forward / backward (one CUDA graph) -> flat NCCL all-reduce -> clip + AdamW + decay kernel + working-copy
    refresh (second graph).  Whole-job images/s, max over ranks."""
    from rgb_no_more_b200 import plan as P, train_step as TS, transforms as TF
    B = 64
    stage = TS.TrainStage(dev, arch="swinv2t", batch=B, world=world, rank=rank, use_graph=not args.no_graph)
    with torch.no_grad():           # the reference zero-initialises the block post-norms (identity blocks): randomise them
        with torch.random.fork_rng(devices=[dev]):
            torch.manual_seed(11997733)
            stage.eng.flat.add_(0.1 * torch.randn_like(stage.eng.flat) * (stage.eng.flat == 0))
    stage.eng.refresh_weights()
    tf = TF.FusedDCT(dev, "train", P.AUGLIST_VITS, 2, 9, torch.bfloat16, out_size=32)
    labels = torch.randint(0, 1000, (B,), device=dev)
    gen = torch.Generator().manual_seed(7 + rank)
    losses = []

    def step(i):
        y, c, q = (t[:B] for t in dev_pool[i % len(dev_pool)])
        plans = tf.sample_plans_packed(B, 64, 64, clamp_in=[False] * B, generator=gen)
        x = tf.run(y, c, q, plans, needs_stats=bool(plans["needs_stats"].any()), out=stage.x_static)
        losses.append(stage.step(x, labels).clone())
    n = max(args.steps, 10)
    ms = timed_loop(step, n, 4) / n
    return {"value": B * world / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "batch_per_gpu": B, "n_gpus": world,
            "loss_first": float(losses[0]), "loss_last": float(losses[-1]),
            "what": "SwinV2-T DCT (window 8) train step: K0 in the Swin layout (RandAugment mix, fresh plans per step) -> mixup -> forward with "
                    "saved activations -> CE -> backward (CUDA graph) -> " + ("flat all-reduce (NCCL) -> " if world > 1 else "") +
                    "clip + AdamW + decoupled decay kernel + bf16 working-copy refresh (CUDA graph); tcgen05 GEMMs, mma.sync window "
                    "attention forward and backward, CUDA-core LayerNorm backward"}


def vitti_configs(args, dev, dev_pool, dev_plans, tf, tf_eval, labels_pool, timed_loop):
    """BASELINE configs 1 and 2 on this arm.  (1) ViT-Ti --domain=DCT, batch 8, 64 synthetic 512x512 JPEGs, eval-only forward: from
    JPEG bytes (host Huffman decode, 1 thread as the reference's single process) -> K0 eval geometry -> ViT-Ti forward, logits
    to the host.  (2) ViT-Ti DCT, batch 256, train step with coefficients resident (K0 + tcgen05 ViT-Ti; the reference's config
    names torch attention, this arm has only its own attention kernels)."""
    from rgb_no_more_b200 import dct_manip as dm, plan as P, synth, train_step as TS, vit as V
    out = {}
    jpegs = synth.synth_jpeg_set(64)
    with torch.random.fork_rng(devices=[dev]):
        torch.manual_seed(11997733)
        m = V.ViT(patch_size=16, emb_size=192, depth=12, n_classes=1000, drop_p=0.0, pixel_space="DCT", ver=1, use_subblock=True,
                  device=dev, num_heads=3, head_size=64).eval()
    host = torch.empty((8, 1000), dtype=torch.float32).pin_memory()
    eplans = tf_eval.sample_plans(8)

    def eval_pass(_i):
        for b in range(0, 64, 8):
            y, c, q, fl = dm.decode_batch(jpegs[b:b + 8], 64, 64, nthreads=1)
            x = tf_eval.run(y.to(dev, non_blocking=True), c.to(dev, non_blocking=True), q.to(dev, non_blocking=True), eplans,
                            clamp_in=fl.tolist())
            with torch.no_grad():
                host.copy_(m(x), non_blocking=True)
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    eval_pass(0)
    reps = 3
    t0 = time.perf_counter()
    for i in range(reps):
        eval_pass(i)
    out["config1_eval_b8_from_jpeg"] = {"value": 64 * reps / (time.perf_counter() - t0), "unit": UNIT,
                                        "what": "ViT-Ti DCT eval forward, batch 8, 64 JPEGs: 1 host decode thread -> H2D -> K0 (eval "
                                                "geometry) -> ViT-Ti forward -> logits D2H, wall clock incl. decode"}
    st = TS.TrainStage(dev, arch="vitti", batch=args.batch, world=1)
    xb = torch.empty((args.batch, 196, 384), dtype=torch.bfloat16, device=dev)

    def step(i):
        k = i % N_POOL
        x = tf.run(*dev_pool[k], None, plans_dev=dev_plans[k], out=st.x_static)
        st.step(x, labels_pool[k])
    n = max(args.steps // 2, 5)
    ms = timed_loop(step, n, 3) / n
    out["config2_train_b256"] = {"value": args.batch / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                                 "what": "ViT-Ti DCT train step (K0 -> fwd/bwd/AdamW), batch 256, coefficients resident"}
    del st, xb
    return out


def swin_eval_forward(args, dev, dev_pool, clamp_flags, timed_loop, host_pool=None):
    from rgb_no_more_b200 import plan as P, swin as S, transforms as TF
    B = args.batch
    torch.manual_seed(11997733)
    model = S.SwinTransformerV2(img_size=256, patch_size=4, embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24],
                                window_size=8, mlp_ratio=4, drop_path_rate=0.2, pretrained_window_sizes=[0, 0, 0, 0],
                                device="cpu", pixel_space="dct")
    with torch.no_grad():       # the reference zero-initialises the block post-norms (identity blocks): randomise them
        for p in model.parameters():
            if p.ndim == 1:
                p.add_(0.1 * torch.randn_like(p))
    model.eval().to(dev)
    eng = model.prepare(dev)
    tf = TF.FusedDCT(dev, "test", None, 0, 0, torch.bfloat16, out_size=32)
    plans = tf.sample_plans(B)
    pdev = [torch.from_numpy(P.pack_plans(plans, fl, out_size=32).view(np.uint8).reshape(B, -1).copy()).to(dev) for fl in clamp_flags]
    x = torch.empty((B, 4096, 24), dtype=torch.bfloat16, device=dev)
    n_pool = len(dev_pool)

    def step(k):
        tf.run(*dev_pool[k], None, plans_dev=pdev[k], out=x)
        return eng.forward(x)
    with torch.no_grad():
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for k in range(2):
                logits = step(k)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        eng.launches = 0
        step(0)
        launches = eng.launches + 2
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for k in range(n_pool):
                step(k)
        n = max(args.steps // 2, 5)
        ms = timed_loop(lambda i: g.replay(), n, 3) / (n * n_pool)
    # end to end through the public API: pinned HOST coefficient buffers in, host logits out, every batch; the H2D copy of
    # batch i+1 runs on a copy stream while batch i computes (one copy per batch inside the timed region)
    e2e = None
    if host_pool is not None:
        h2d = torch.cuda.Stream(device=dev)
        stage = [tuple(torch.empty_like(t, device=dev) for t in host_pool[0]) for _ in range(2)]
        ev_ready = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        out_host = torch.empty((B, 1000), dtype=torch.float32).pin_memory()
        copied = set()

        def issue_copy(i):
            sidx = i % 2
            with torch.cuda.stream(h2d):
                h2d.wait_event(ev_free[sidx])
                for dst, src in zip(stage[sidx], host_pool[i % n_pool]):
                    dst.copy_(src, non_blocking=True)
                ev_ready[sidx].record(h2d)
            copied.add(i)

        def step_e2e(i):
            sidx = i % 2
            if i not in copied:
                issue_copy(i)
            torch.cuda.current_stream().wait_event(ev_ready[sidx])
            with torch.no_grad():
                tf.run(*stage[sidx], None, plans_dev=pdev[i % n_pool], out=x)
                lg = model(x)                                  # the call a user makes: (B, 4096, 24) -> logits
            ev_free[sidx].record()
            issue_copy(i + 1)
            out_host.copy_(lg, non_blocking=True)
            copied.discard(i)
        for evf in ev_free:
            evf.record()
        n_e = max(args.steps // 2, 5)
        ms_e = timed_loop(step_e2e, n_e, 3) / n_e
        torch.cuda.synchronize()
        e2e = {"value": B / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(sum(t.numel() * t.element_size() for t in host_pool[0])),
               "d2h_bytes_per_step": B * 1000 * 4}
    fl = 2 * 4096 * 24 * 96 + 2 * 768 * 1000            # algorithmic forward FLOPs per image: contractions + attention MACs x 2
    for s_, depth in enumerate((2, 2, 6, 2)):
        T, Cd = 4096 // 4 ** s_, 96 * 2 ** s_
        fl += depth * (24 * T * Cd * Cd + 256 * T * Cd) + (4 * T * Cd * Cd if s_ < 3 else 0)
    return {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_batch": ms, "batch": B, "gflop_per_image_fwd": fl / 1e9,
            "achieved_tflops": fl * B / ms / 1e9, "gpu_launches_per_batch": launches, "logits_finite": bool(torch.isfinite(logits).all()),
            "e2e": e2e,
            "what": "SwinV2-T DCT (window 8, 256 px) eval forward: K0 in the Swin layout (Resize_DCT(32)) + SwinTransformerV2 forward, "
                    "no gradients, coefficients resident in HBM, CUDA-graph replay"}


# ----------------------------------------------------------------------------------------------
# CPU legs (oracle port of the reference path) -- the only place bench.py touches oracle/
#
# The reference's own pipeline for this workload is: DataLoader worker processes running the DCT transforms one image
# at a time with one torch thread each (datasets.py:542-556, pipeline_utils.py:125), then the model's training step
# (train.py:146-176) in the main process with every core.  The CPU arm restates exactly that on a bounded sample:
# oracle/dct_oracle.transform_embed in `cores` forked workers -> mixup -> oracle/vit_oracle.train_step (ViT of the
# benchmarked arch, fp32) on all cores.  /root/reference itself cannot be installed.
# ----------------------------------------------------------------------------------------------
def _cpu_transform_worker(job):
    from oracle import dct_oracle as O
    y, c, q, plan, filters = job
    return O.transform_embed(y, c, q, plan, filters)


def _cpu_jobs(n_images: int):
    from rgb_no_more_b200 import plan as P, synth
    bank = P.FilterBank()
    torch.manual_seed(11997733)
    y, c, q = synth.synth_coefficients(n_images, 64, 64, seed=synth.SEED)
    jobs = []
    for b in range(n_images):
        pl = P.sample_train_plan(64, 64, P.AUGLIST_VITS, 2, 9, bank)
        jobs.append((torch.from_numpy(y[b]).reshape(1, 64, 64, 8, 8), torch.from_numpy(c[b]).reshape(2, 32, 32, 8, 8),
                     torch.from_numpy(q[b]).reshape(3, 8, 8), pl, bank.table))
    return jobs


def _cpu_chunk(jobs):
    torch.set_num_threads(1)
    return torch.stack([_cpu_transform_worker(j) for j in jobs]) if jobs else torch.zeros((0, 196, 384))


def _cpu_state_dict(arch: str):
    from rgb_no_more_b200 import train_step as TS, vit as V
    cfg = TS.ARCHS[arch]
    torch.manual_seed(11997733)
    m = V.ViT(patch_size=16, emb_size=cfg["emb_size"], depth=cfg["depth"], n_classes=1000, drop_p=0.0, pixel_space="DCT",
              ver=1, use_subblock=True, device="cpu", num_heads=cfg["num_heads"], head_size=64)     # parameters only, never run
    return {k: v.detach().clone() for k, v in m.state_dict().items()}, cfg["wd"]


def cpu_train_steps(args, n_images: int, cores: int, steps: int, warmup: int):
    """`warmup + steps` CPU training steps over the same `n_images`-image sample; returns seconds per timed step
    and the split (data path, model)."""
    import multiprocessing as mp
    from oracle import vit_oracle as VO
    jobs = _cpu_jobs(n_images)
    pool = mp.get_context("fork").Pool(cores) if cores > 1 else None      # forked before the parent spins up its torch threads
    try:
        chunks = [jobs[i::cores] for i in range(cores)]
        sd, wd = _cpu_state_dict(args.arch)
        torch.set_num_threads(cores)
        state, labels = {}, torch.arange(n_images) % 1000
        gen = torch.Generator().manual_seed(5)
        t_data, t_model = [], []
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            parts = pool.map(_cpu_chunk, chunks) if pool is not None else [_cpu_chunk(jobs)]
            x = torch.cat([p_ for p_ in parts if p_.shape[0]])
            t1 = time.perf_counter()
            lam = torch._sample_dirichlet(torch.tensor([0.2, 0.2]), generator=gen).sort(descending=True)[0]
            onehot = torch.nn.functional.one_hot(labels, 1000).float()
            x = lam[0] * x + lam[1] * x.roll(1, 0)                                        # RandomMixup_DCT (cls_transforms.py:135-182)
            soft = lam[0] * onehot + lam[1] * onehot.roll(1, 0)
            VO.train_step(sd, x.reshape(n_images, 14, 14, 384), soft, state, lr=3e-3 * (s + 1) / 10000, wd=wd)
            t2 = time.perf_counter()
            if s >= warmup:
                t_data.append(t1 - t0)
                t_model.append(t2 - t1)
        return float(np.mean(t_data) + np.mean(t_model)), float(np.mean(t_data)), float(np.mean(t_model))
    finally:
        if pool is not None:
            pool.close()
            pool.join()


def _cpu_desc(args, n, cores, t_data, t_model):
    return (f"{n} images per step of the same workload ({args.arch} DCT train step, RandAugment num_ops=2 magnitude=9): "
            f"oracle/dct_oracle.py data path in {cores} worker processes ({t_data * 1e3:.0f} ms) + oracle/vit_oracle.py "
            f"fp32 forward/backward/AdamW on {cores} torch threads ({t_model * 1e3:.0f} ms)")


def cpu_baseline(args):
    """Our arm's `cpu_baseline`: the reference arm below, run once in a fresh process (this one already holds a CUDA
    context and a warm torch thread pool -- forking it is not safe) on a smaller sample."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1", "--arch", args.arch,
           "--cpu-sample", str(args.cpu_sample if args.cpu_sample > 0 else 64), "--cpu-budget-s", "30"]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
    for line in reversed(out.stdout.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)["cpu_baseline"]
    return {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "cpu leg failed: " + out.stderr[-300:]}


def run_reference(args):
    """Reference arm.  With oracle/_ref staged (oracle/make_ref.py, run by build() in the container that has /root/reference)
    this is the REFERENCE ITSELF on the host cores: its DataLoader workers (decode -> dequantise -> DCT transforms), its
    RandomMixup_DCT, its pvit.ViT, its optimiser objects (oracle/ref_arm.py lists what is its and what is ours); same arch,
    same RandAugment setting and -- when `steps + warmup` such steps fit ~4 minutes on this host -- the same 256-image batch
    as our arm.  Without oracle/_ref: the oracle port, as in round 1."""
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    from oracle import ref_arm as RA
    if RA.available() and not args.port:
        want = args.cpu_sample if args.cpu_sample > 0 else args.batch
        n = RA.calibrate_images_per_step(args.arch, cores, want, args.steps + args.warmup, args.cpu_budget_s)
        r = RA.cpu_train_arm(args.arch, n, args.steps, args.warmup, cores)
        sec, kind = r["sec_per_step"], "reference"
        sample = (f"{n} images per step ({'the full batch of our arm' if n == args.batch else f'bounded sample; our arm: {args.batch}'}) "
                  f"of the same workload ({args.arch} DCT train step, RandAugment num_ops=2 magnitude=9): the reference's own "
                  f"datasets.dataset_selector DataLoader with {cores} worker processes (JPEG files -> dct_manip.read_coefficients [this "
                  f"repo's host decoder: the reference's needs jpeglib.h] -> dequantise -> get_transform('imagenet_dct','train')), "
                  f"RandomMixup_DCT, models.plainvit.ViT + CrossEntropyLoss + clip + AdamW + WeightDecay in fp32 (reference default "
                  f"AMP off) on {cores} torch threads; loader wait {r['sec_loader_wait'] * 1e3:.0f} ms + model step {r['sec_model'] * 1e3:.0f} ms")
        workload = (f"reference CPU path (oracle/_ref: the reference's own Python) of the {args.arch} DCT train step from JPEG files, "
                    f"{n} images/step, RandAugment num_ops=2 magnitude=9")
        same = n == args.batch
    else:
        n = args.cpu_sample if args.cpu_sample > 0 else 32
        sec, t_data, t_model = cpu_train_steps(args, n, cores, args.steps, args.warmup)
        kind, sample, same = "port", _cpu_desc(args, n, cores, t_data, t_model), False
        workload = (f"reference CPU path (oracle port) of the {args.arch} DCT train step, bounded sample of {n} "
                    f"images/step (our arm: batch {args.batch}/GPU), RandAugment num_ops=2 magnitude=9")
    val = n / sec
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "cpu_model": _cpu_model(), "same_config": same, "images_per_step": n},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_torch_b200(args):
    """PyTorch-on-B200 baseline: the reference's own model classes under eager PyTorch."""
    from oracle import ref_arm as RA
    if not RA.available():
        print(json.dumps({"impl": "torch_b200", "unavailable": "oracle/_ref not staged"}))
        return
    out = RA.torch_b200_arm(args.arch, args.batch, max(3, args.steps), max(3, args.warmup))
    out["impl"] = "torch_b200"
    print(json.dumps(out))


def torch_b200_baseline(args):
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "torch_b200", "--steps", "5", "--warmup", "3", "--arch", args.arch,
           "--batch", str(args.batch)]
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
        for line in reversed(out.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return {"error": out.stderr[-300:]}
    except Exception as ex:  # noqa: BLE001 -- a side measurement must never cost the headline line
        return {"error": f"{type(ex).__name__}: {ex}"[:300]}


def _cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "torch_b200"])
    ap.add_argument("--port", action="store_true", help="reference arm: force the oracle port even if oracle/_ref is staged")
    ap.add_argument("--cpu-budget-s", type=float, default=240.0, help="reference arm: wall-clock budget of all CPU steps")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--arch", default="vits")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--stage", default="auto", choices=["auto", "k0", "train"])
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly (for ncu launch lists)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="images per CPU step of the reference arm (0: the full batch if it fits the budget)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-swin", action="store_true", help="skip the SwinV2-T eval-forward side measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.stage == "auto":
        args.stage = "train" if os.path.exists(os.path.join(ROOT, "rgb_no_more_b200", "train_step.py")) else "k0"
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch_b200":
        run_torch_b200(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
