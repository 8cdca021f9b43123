#!/usr/bin/env python
"""Headline benchmark: SAM ViT-H ImageEncoderViT forward, img/s at 1024^2 (BASELINE.json).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm on the host cores (oracle port)

A "step" is one forward of the encoder over a batch of 32 synthetic images per GPU (BASELINE.json
configs[1]), called the way the cascade calls it (``image_encoder(x, interm=True)``,
models/sam_maskdecoder_edge.py:337).  One JSON line is printed by rank 0:

  value      whole-job img/s with the inputs already resident in HBM (CUDA events, max over ranks)
  e2e        same metric through the public nn.Module call with HOST buffers: pinned H2D of the batch,
             forward, (N>1: NCCL all-gather of the embeddings), D2H of the embeddings, every step
  roofline   the dominant kernel (the tcgen05 GEMM): algorithmic FLOPs of its launches / their CUDA-event
             time, against the measured dense-bf16 peak in MEASURED_PEAKS.json
  cpu_baseline  the oracle (CPU restatement of the reference) timed on this box's host cores, 1 image
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "vit_h_image_encoder_throughput_1024px"
UNIT = "img/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--model", default="vit_h")
    ap.add_argument("--no-interm", action="store_true", help="call forward(x) instead of forward(x, interm=True)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-decoder", action="store_true", help="skip the mask-decoder / cascade / CLIP-tower legs (SURVEY 8 f1, f2)")
    ap.add_argument("--skip-torch-gpu-baseline", action="store_true",
                    help="skip the leg that times the reference module itself (eager PyTorch) on this GPU")
    ap.add_argument("--skip-train", action="store_true", help="skip the adapter-training step leg (SURVEY 8 f4)")
    ap.add_argument("--total-batch", type=int, default=0,
                    help="strong scaling (BASELINE.json configs[2]): this many images in total, split over the ranks and "
                         "processed in micro-batches of --batch; 0 = weak scaling (--batch images per GPU per step)")
    ap.add_argument("--gather-dtype", choices=["fp32", "bf16"], default="fp32",
                    help="payload of the N>1 embedding all-gather (bf16 halves the bytes, adds ~1.1e-3 rel-L2)")
    ap.add_argument("--e2e-chunk", type=int, default=16,
                    help="images per chunk of the host-buffer call timed as e2e (upload / encoder / download pipeline)")
    return ap.parse_args()


def workload_config(args, cfg, world):
    per_rank = args.total_batch // world if args.total_batch else args.batch
    return {
        "workload": f"SAM {args.model} ImageEncoderViT forward (interm={not args.no_interm}), bf16, "
                    + (f"{args.total_batch} images per step in total ({per_rank} per GPU in micro-batches of {args.batch}), "
                       if args.total_batch else f"batch {args.batch} per GPU, ")
                    + f"{cfg.img_size}x{cfg.img_size} synthetic images, random-init weights "
                    "(BASELINE.json configs[1]; configs[2] sharding + embedding all-gather at N>1)",
        "batch_per_gpu": per_rank,
        "micro_batch": args.batch,
        "global_batch": per_rank * world,
        "img_size": cfg.img_size,
        "parallelism": f"dp{world} (batch-sharded, weights replicated)",
        "l2": "inputs larger than L2: 12.6 MB/image input, ~85 MB/image of activations streamed per block",
    }


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
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
        sm, smax, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            p = json.load(open(path))
            return {"burst": float(p["bf16_tflops"]), "sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    "hbm_gbs": float(p["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
        except Exception:
            pass
    return {"burst": 1590.0, "sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def gemm_flops_per_image(cfg):
    """Algorithmic FLOPs of the launches of the tcgen05 GEMM kernel (everything in cfg.flops_per_image()
    except attention and the adapter's SIMT embedding kernel)."""
    n, d, r = cfg.tokens, cfg.embed_dim, cfg.rank
    kp = cfg.in_chans * cfg.patch_size ** 2
    lin = cfg.depth * (2 * n * d * 3 * d + 2 * n * d * d + 2 * 2 * n * d * cfg.mlp_dim)
    stem = 2 * n * kp * d + 2 * n * kp * r
    adapter = cfg.depth * (2 * n * r * r + 2 * n * r * d)
    neck = 2 * n * d * cfg.out_chans + 2 * n * cfg.out_chans * cfg.out_chans * 9
    return float(lin + stem + adapter + neck)


# ----------------------------------------------------------------------------------------------
def cpu_reference_forward(cfg, sd_cpu):
    """The reference's CPU implementation of the path as a callable x -> (embedding, interm list), and its kind:
    "reference" = the reference's own ImageEncoderViT module from oracle/_ref/refsam (staged, unmodified, by
    oracle/stage_ref.py), "port" = the oracle restatement when nothing is staged."""
    from oracle import ref_import
    if ref_import.staged_encoder_available():
        ref = ref_import.build_reference_encoder(cfg, staged=True).eval()
        ref.load_state_dict(sd_cpu, strict=True)
        return (lambda x: ref(x, interm=True)), "reference"
    from oracle import encoder_oracle as orc
    return (lambda x: orc.encoder_forward(sd_cpu, cfg, x, interm=True)), "port"


def time_reference_cpu(cfg, sd_cpu, steps, warmup, budget_s=240.0):
    """One image per step on all host threads.  Returns a dict: img/s, seconds per step, steps / warm-ups actually run
    (capped to the budget), the kind of implementation timed, and the last embedding (for the parity line)."""
    import torch
    from camouflaged_vlm_b200.synth import seeded_images
    torch.set_num_threads(os.cpu_count() or 1)
    fwd, kind = cpu_reference_forward(cfg, sd_cpu)
    x = seeded_images(cfg, 1, 0)
    t0 = time.perf_counter()
    out = fwd(x)[0]                                           # first pass doubles as warm-up / calibration
    first = time.perf_counter() - t0
    warm_done = 1
    while warm_done < warmup and (warm_done + 1 + max(1, steps)) * first < budget_s:
        fwd(x)
        warm_done += 1
    steps_run = max(1, min(steps, int((budget_s - warm_done * first) / max(first, 1e-9))))
    times = []
    for _ in range(steps_run):
        t0 = time.perf_counter()
        out = fwd(x)[0]
        times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return {"ips": 1.0 / dt, "dt": dt, "best_dt": min(times), "steps": steps_run, "warmup": warm_done, "kind": kind,
            "out": out, "image": "seeded_images(cfg, 1, 0)[0]"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores -- the reference
    module itself from oracle/_ref/refsam when staged (oracle/stage_ref.py), else the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import torch
    from camouflaged_vlm_b200 import PRESETS
    from camouflaged_vlm_b200.synth import seeded_state_dict
    cfg = PRESETS[args.model]
    sd = seeded_state_dict(cfg, 0)
    with torch.no_grad():
        r = time_reference_cpu(cfg, sd, args.steps, max(args.warmup, 1))
    cores = os.cpu_count() or 1
    what = ("the reference's ImageEncoderViT module (oracle/_ref/refsam, unmodified)" if r["kind"] == "reference"
            else "oracle port of the reference (nothing staged under oracle/_ref)")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["ips"], "unit": UNIT, "n_gpus": world, "steps": r["steps"],
        "warmup": r["warmup"], "ms_per_step": r["dt"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, cfg, world),
        "cpu_baseline": {"value": r["ips"], "unit": UNIT, "cores": cores, "kind": r["kind"],
                         "sample": f"{what}, fp32, torch {torch.__version__}, {cores} threads: 1 image per step (a batch-32 "
                                   f"step is 32 independent images; the CPU rate is per image), {r['steps']} timed steps "
                                   f"after {r['warmup']} warm-up (requested {args.steps}/{args.warmup}, capped to a "
                                   f"4-minute budget), best step {r['best_dt']:.2f} s"},
        "e2e": {"value": r["ips"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def time_torch_gpu_baseline(cfg, sd_cpu, dev, ours, batches=(16, 8, 4)):
    """SURVEY 2.2 / BASELINE.md: "the bar on B200 is eager PyTorch + cuBLAS on the same box".  The reference module
    itself (oracle/_ref/refsam) moved to the GPU, stock code path: fp32 eager and under torch.autocast(bf16), at the
    largest batch of `batches` that fits (it materialises 32 prompt tensors and the [B*16, 4096, 4096] fp32 scores),
    next to this repository's encoder on the same images.  A baseline, never the product."""
    import torch
    from camouflaged_vlm_b200.synth import seeded_images
    from oracle import ref_import
    from oracle.encoder_oracle import rel_l2
    if not ref_import.staged_encoder_available():
        return {"unavailable": "no staged reference under oracle/_ref (run oracle/stage_ref.py in the container)"}
    ref = ref_import.build_reference_encoder(cfg, staged=True).eval()
    ref.load_state_dict(sd_cpu, strict=True)
    ref = ref.to(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timeit(fn, iters):
        fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    rep = None
    for B in batches:
        try:
            x = seeded_images(cfg, B, 0).to(dev)
            with torch.no_grad():
                y32 = ref(x, interm=True)[0]
                ms32 = timeit(lambda: ref(x, interm=True), 2)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    y16 = ref(x, interm=True)[0].float()
                    ms16 = timeit(lambda: ref(x, interm=True), 3)
                yo = ours(x)
                mso = timeit(lambda: ours(x), 5)
            rep = {"batch": B, "what": "the reference ImageEncoderViT module (oracle/_ref/refsam, unmodified) on this GPU, "
                                       "forward(x, interm=True), same weights and images",
                   "torch_fp32_eager": {"img_per_s": B / ms32 * 1e3, "ms": ms32},
                   "torch_autocast_bf16": {"img_per_s": B / ms16 * 1e3, "ms": ms16, "rel_l2_vs_fp32": rel_l2(y16, y32)},
                   "ours_same_batch": {"img_per_s": B / mso * 1e3, "ms": mso, "rel_l2_vs_fp32": rel_l2(yo, y32)}}
            break
        except torch.cuda.OutOfMemoryError:
            torch.cuda.empty_cache()
            continue
    del ref
    torch.cuda.empty_cache()
    return rep or {"unavailable": f"out of memory at every batch of {batches}"}


# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from camouflaged_vlm_b200 import PRESETS, _lib, build_encoder
    from camouflaged_vlm_b200.dist import ChunkGatherer
    from camouflaged_vlm_b200.synth import param_shapes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = PRESETS[args.model]
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    interm = not args.no_interm
    # weak scaling (default): B images per GPU per step.  Strong scaling (--total-batch T, BASELINE.json configs[2]):
    # T images per step in total, T / world per rank, encoded in micro-batches of B.
    if args.total_batch:
        if args.total_batch % (world * B):
            raise SystemExit(f"--total-batch {args.total_batch} must be a multiple of world * batch = {world * B}")
        per_rank = args.total_batch // world
    else:
        per_rank = B
    n_micro = per_rank // B
    gdtype = torch.bfloat16 if args.gather_dtype == "bf16" else torch.float32

    # random-init weights of the architecture (same recipe as synth.seeded_state_dict, generated on the GPU)
    model = build_encoder(cfg).eval().to(dev)
    gen = torch.Generator(device=dev).manual_seed(1234)
    with torch.no_grad():
        for name, p in model.named_parameters():
            p.copy_(torch.randn(p.shape, generator=gen, device=dev) * 0.02)
            if name.endswith(("norm1.weight", "norm2.weight", "neck.1.weight", "neck.3.weight")):
                p.add_(1.0)
    assert set(dict(model.named_parameters())) == set(param_shapes(cfg))
    x_dev = torch.randn(B, cfg.in_chans, cfg.img_size, cfg.img_size, generator=gen, device=dev)
    tail = (cfg.out_chans, cfg.grid, cfg.grid)
    # N > 1: the embeddings of every micro-batch are all-gathered on a side stream underneath the encoder of the next
    # one (two result buffers alternate), and the timed region ends when the last gather has landed -- the exchange
    # step of configs[2] is INSIDE `value`
    gatherers = [ChunkGatherer(B, tail, dev, gdtype) for _ in range(2)] if world > 1 else None
    micro_done = [0]

    def step(x):
        for _ in range(n_micro):
            with torch.no_grad():
                res = model(x, interm=interm)
            if gatherers is not None:
                emb = res[0] if interm else res
                g = gatherers[micro_done[0] % 2]
                g.submit(0, B, emb)
                micro_done[0] += 1

    def drain():
        if gatherers is not None:
            for g in gatherers:
                g.finish()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(W):
        step(x_dev)
    drain()
    barrier()

    # ---- timed region: inputs resident in HBM
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    e_enc = torch.cuda.Event(enable_timing=True)
    for _ in range(K):
        step(x_dev)
    e_enc.record()          # this rank's encoders are enqueued; the gathers run on the side stream
    drain()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    by_rank = None
    if world > 1:
        # where the scaling efficiency goes: every rank's own encoder time (GPU-to-GPU spread, SMs lent to NCCL) next to
        # its time including the wait for the last gather
        t = torch.tensor([e0.elapsed_time(e_enc) / K, e0.elapsed_time(e1) / K], device=dev, dtype=torch.float64)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        by_rank = {"encoder_ms_per_step": [round(float(a[0]), 3) for a in allt],
                   "with_gather_ms_per_step": [round(float(a[1]), 3) for a in allt]}
    launches = model.last_launch_count * K * n_micro
    value = world * per_rank * K / (ms_total * 1e-3)
    gather_info = None
    if gatherers is not None:
        gather_info = {"inside_value": True, "dtype": args.gather_dtype, "bytes_received_per_rank_per_step":
                       (world - 1) * per_rank * tail[0] * tail[1] * tail[2] * (2 if gdtype == torch.bfloat16 else 4),
                       "how": "dist.ChunkGatherer: all_gather of each micro-batch's embeddings on a side stream under the "
                              "next micro-batch's encoder; the timed region waits for the last one"}
        gather_info["by_rank"] = by_rank
    gatherers = None

    def plain_step(x):
        with torch.no_grad():
            return model(x, interm=interm)

    # ---- per-kernel CUDA-event timing over the same K steps (events bracket every launch on the stream)
    lib = _lib.load()
    lib.cvlm_encoder_set_profiling(model._handle, 1)
    for _ in range(K):
        plain_step(x_dev)
    torch.cuda.synchronize()
    import ctypes as C
    ms_kind = (C.c_float * len(_lib.PROF_KINDS))()
    n_kind = (C.c_int * len(_lib.PROF_KINDS))()
    _lib.check(lib.cvlm_encoder_get_profile(model._handle, ms_kind, n_kind), "get_profile")
    lib.cvlm_encoder_set_profiling(model._handle, 0)
    kernels = {k: {"ms_per_step": ms_kind[i] / K, "launches_per_step": n_kind[i] // K} for i, k in enumerate(_lib.PROF_KINDS)}
    gemm_ms = sum(kernels[k]["ms_per_step"] for k in _lib.PROF_KINDS if k.startswith("gemm"))
    gemm_launches = sum(kernels[k]["launches_per_step"] for k in _lib.PROF_KINDS if k.startswith("gemm"))
    peaks = load_peaks()
    gemm_tflops = B * gemm_flops_per_image(cfg) / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    step_tflops = per_rank * cfg.flops_per_image() / (ms_total / K * 1e-3) / 1e12
    # DRAM traffic per GEMM launch from the committed `ncu --set full` capture of this kernel (same batch)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_gemm_latest.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("batch") == B and args.model == "vit_h":
                traffic = tj.get("mean_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {
        "bound": "tensor", "kernel": "gemm_bf16_tcgen05_kernel", "achieved": gemm_tflops, "peak": peaks["sustained"],
        "unit": "TFLOP/s", "frac": gemm_tflops / peaks["sustained"], "traffic": traffic,
        "traffic_note": "mean dram__bytes_read+write per launch over the qkv/proj/lin1/lin2 launches of one block, "
                        "ncu --set full at the same batch (profiles/ncu_gemm_latest.json); algorithmic bytes per "
                        "launch are in DESIGN.md section 5",
        "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
        "launches_per_step": gemm_launches, "avg_launch_ms": gemm_ms / max(gemm_launches, 1),
        "share_of_step": gemm_ms * n_micro / (ms_total / K),
        "whole_step": {"achieved_tflops": step_tflops, "frac_of_sustained": step_tflops / peaks["sustained"],
                       "frac_of_burst": step_tflops / peaks["burst"], "frac_of_nominal_2250": step_tflops / 2250.0,
                       "flops_per_image": cfg.flops_per_image()},
    }

    # ---- end to end through the public call with host buffers
    e2e = None
    if not args.skip_e2e:
        x_host = torch.empty(B, cfg.in_chans, cfg.img_size, cfg.img_size, pin_memory=True)
        x_host.copy_(x_dev)
        out_host = torch.empty(B, cfg.out_chans, cfg.grid, cfg.grid, pin_memory=True)
        CH = max(1, min(args.e2e_chunk, B))
        e2e_gather = [ChunkGatherer(B, tail, dev, gdtype) for _ in range(2)] if world > 1 else None
        e2e_calls = [0]

        def e2e_step():
            # the public host-buffer call: chunked H2D on a copy stream under the encoder, D2H of this rank's
            # embeddings per chunk; at N > 1 each chunk's embeddings are also all-gathered on the devices (configs[2])
            # on a side stream under the next chunk's encoder
            for _ in range(n_micro):
                g = e2e_gather[e2e_calls[0] % 2] if e2e_gather is not None else None
                e2e_calls[0] += 1
                with torch.no_grad():
                    model.forward_host(x_host, out_host=out_host, chunk=CH, interm=interm,
                                       on_chunk=g.submit if g is not None else None)

        def e2e_drain():
            if e2e_gather is not None:
                for g in e2e_gather:
                    g.finish()

        e2e_step()
        e2e_drain()
        barrier()
        e0.record()
        for _ in range(K):
            e2e_step()
        e2e_drain()
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        e2e = {"value": world * per_rank * K / (ms_e2e * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": world * n_micro * x_host.numel() * 4,
               "d2h_bytes_per_step": world * n_micro * out_host.numel() * 4, "ms_per_step": ms_e2e / K,
               "note": f"ImageEncoderViT.forward_host: pinned host -> device in chunks of {CH} on a copy stream under "
                       "forward(x, interm) of the previous chunk, "
                       + (f"all-gather ({args.gather_dtype}) of each chunk's embeddings on a side stream under the next "
                          "chunk's encoder (dist.ChunkGatherer), " if world > 1 else "")
                       + "device -> pinned host copy of each rank's embeddings (per chunk, same copy stream); whole-job bytes"}
        e2e_gather = None

    # ---- the next row of the hot path: the mask decoder on this step's embeddings, and the
    # encoder -> decoder -> postprocess cascade; reported beside the headline, not part of it
    f1 = None
    if rank == 0 and world == 1 and not args.skip_decoder and cfg.out_chans == 256 and cfg.grid % 8 == 0:
        from camouflaged_vlm_b200 import mask_decoder as md
        dec = md.build_mask_decoder(seed=0).to(dev)
        gen = torch.Generator().manual_seed(1)
        pe = torch.randn(1, 256, cfg.grid, cfg.grid, generator=gen).to(dev)
        sparse = torch.randn(B, 2, 256, generator=gen).to(dev)
        dense = (torch.randn(1, 256, 1, 1, generator=gen) * 0.05).to(dev).expand(B, -1, cfg.grid, cfg.grid)
        with torch.no_grad():
            emb = model(x_dev)

        def decode(e):
            with torch.no_grad():
                m, _, _ = dec(e, [], pe, sparse, dense, False)
                return md.postprocess_masks(m, cfg.img_size, cfg.img_size, cfg.img_size)

        def cascade():
            with torch.no_grad():
                f, _ = model(x_dev, interm=True)
            return decode(f)

        res = {}
        for name, fn, n in (("decoder", lambda: decode(emb), 10), ("cascade", cascade, max(3, K // 2))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / n
        f1 = {"metric": "mask_decoder_img_per_s", "value": B / res["decoder"] * 1e3, "unit": UNIT,
              "ms_per_step": res["decoder"], "gpu_launches": int(dec.last_launch_count) + 1,
              "workload": f"MaskDecoder_Edge + postprocess_masks on {B} embeddings [256,{cfg.grid},{cfg.grid}] -> "
                          f"{cfg.img_size}x{cfg.img_size} logits, 2 prompt tokens per image, random-init weights",
              "cascade": {"value": B / res["cascade"] * 1e3, "unit": UNIT, "ms_per_step": res["cascade"],
                          "what": "ImageEncoderViT.forward(x, interm=True) -> MaskDecoder_Edge -> postprocess_masks, inputs resident"}}
        del dec
        torch.cuda.empty_cache()

    # ---- the Alpha-CLIP ViT-L/14@336 visual tower with MaPLe prompts (the cascade runs it twice per
    # image); reported beside the headline like the decoder
    f2 = None
    if rank == 0 and world == 1 and not args.skip_decoder:
        from camouflaged_vlm_b200 import alpha_clip as ac
        tw = ac.VisionTransformer_MaPLe(336, 14, 1024, 24, 16, 768, design_details={"trainer": "MaPLe", "maple_length": 2}).eval()
        gen = torch.Generator().manual_seed(2)
        with torch.no_grad():
            for n, p in tw.named_parameters():
                std = 0.5 if n.startswith("conv1") else 1024 ** -0.5
                p.copy_(torch.randn(p.shape, generator=gen) * std + (1.0 if (p.dim() == 1 and n.endswith("weight")) else 0.0))
        tw = tw.to(dev)
        timg = torch.randn(B, 3, 336, 336, generator=gen).to(dev)
        talpha = (torch.rand(B, 1, 336, 336, generator=gen) > 0.5).float().to(dev)
        tshared = (torch.randn(2, 1024, generator=gen) * 0.02).to(dev)
        tdeep = [(torch.randn(2, 1024, generator=gen) * 0.02).to(dev) for _ in range(8)]
        with torch.no_grad():
            for _ in range(3):
                tw(timg, talpha, tshared, tdeep)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                tw(timg, talpha, tshared, tdeep)
            e1.record()
            torch.cuda.synchronize()
        tms = e0.elapsed_time(e1) / 10
        tflop = B * (24 * (2.0 * 579 * 1024 * 12 * 1024 + 4.0 * 579 * 579 * 1024) + 2.0 * 576 * 1024 * 784) / (tms * 1e-3) / 1e12
        f2 = {"metric": "alpha_clip_vit_l14_336_tower_img_per_s", "value": B / tms * 1e3, "unit": UNIT, "ms_per_step": tms,
              "gpu_launches": int(tw.last_launch_count), "achieved_tflops": tflop, "frac_of_sustained": tflop / load_peaks()["sustained"],
              "workload": f"VisionTransformer_MaPLe ViT-L/14@336 on {B} images, 579 tokens (2 context tokens, 8 deeper prompts), "
                          "one pass (the cascade runs two per image), random-init weights"}
        del tw
        torch.cuda.empty_cache()

    # ---- one adapter-only training step (the reference's training configuration: every encoder parameter
    # frozen except prompt_generator.*, train_ovcos_maskdecoder_edge.py:301-303) -- training forward that keeps the
    # activations + backward pass to the 70 adapter gradients, through loss.backward() on the nn.Module
    f4 = None
    if rank == 0 and world == 1 and not args.skip_train:
        tb = min(8, B)
        try:
            for n, p in model.named_parameters():
                p.requires_grad_("prompt_generator" in n)
            model.train()
            xt = x_dev[:tb]
            cot = torch.randn(tb, cfg.out_chans, cfg.grid, cfg.grid, device=dev)
            adapter = [p for p in model.parameters() if p.requires_grad]

            def train_step():
                for p in adapter:
                    p.grad = None
                o, _ = model(xt, interm=True)
                (o * cot).sum().backward()

            for _ in range(2):
                train_step()
            torch.cuda.synchronize()
            n_t = 4
            e0.record()
            for _ in range(n_t):
                train_step()
            e1.record()
            torch.cuda.synchronize()
            t_ms = e0.elapsed_time(e1) / n_t
            with torch.no_grad():
                model.eval()
                for _ in range(2):
                    model(xt, interm=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(n_t):
                    model(xt, interm=True)
                e1.record()
                torch.cuda.synchronize()
            f_ms = e0.elapsed_time(e1) / n_t
            saved_gb = _lib.check(lib.cvlm_encoder_saved_bytes(model._handle, tb)) / 1e9
            f4 = {"metric": "adapter_train_step_img_per_s", "value": tb / t_ms * 1e3, "unit": UNIT, "ms_per_step": t_ms,
                  "batch": tb, "inference_forward_ms_same_batch": f_ms, "step_over_forward": t_ms / f_ms,
                  "gpu_launches": int(model.last_launch_count + model.last_backward_launch_count),
                  "saved_activations_gb": saved_gb, "grads_finite": bool(all(torch.isfinite(p.grad).all() for p in adapter)),
                  "workload": f"forward_train + backward of {tb} ViT-H images (gradients of the {len(adapter)} "
                              "prompt_generator tensors, trunk frozen), loss = <embedding, fixed cotangent>"}
        except torch.cuda.OutOfMemoryError:
            f4 = {"unavailable": "out of memory"}
        finally:
            model.eval()
            for p in model.parameters():
                p.requires_grad_(True)
            model._workspace_bwd = None
            torch.cuda.empty_cache()

    # ---- CPU baseline beside it (rank 0, single-GPU run only): the reference module itself (oracle/_ref/refsam) when
    # staged, else the oracle port, with THIS run's weights; its embedding of the timed image is kept and compared with
    # the CUDA encoder's on the same image -> `parity` (north_star: rel-L2 <= 1e-2)
    cpu = None
    parity = None
    tgb = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        from camouflaged_vlm_b200.synth import seeded_images
        from oracle.encoder_oracle import rel_l2
        sd_cpu = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
        with torch.no_grad():
            r = time_reference_cpu(cfg, sd_cpu, 3, 1, budget_s=150.0)
            ours1 = model(seeded_images(cfg, 1, 0).to(dev)).cpu()
        what = ("the reference ImageEncoderViT module (oracle/_ref/refsam, unmodified)" if r["kind"] == "reference"
                else "oracle port of the reference")
        cpu = {"value": r["ips"], "unit": UNIT, "cores": os.cpu_count() or 1, "kind": r["kind"],
               "sample": f"{what}, fp32, torch {torch.__version__}, on 1 image of the same workload with this run's "
                         f"weights: {r['steps']} timed passes after {r['warmup']} warm-up, mean {r['dt']:.2f} s "
                         f"(best {r['best_dt']:.2f} s) per image"}
        parity = {"rel_l2": rel_l2(ours1, r["out"]), "gate": 1e-2, "image": r["image"], "weights_seed": 1234,
                  "against": r["kind"] + " (CPU fp32) embedding of the same image with the same weights"}
        if not args.skip_torch_gpu_baseline and args.model == "vit_h":
            def ours_fn(x):
                return model(x, interm=True)[0]
            tgb = time_torch_gpu_baseline(cfg, sd_cpu, dev, ours_fn)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong" if args.total_batch else "weak",
            "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": workload_config(args, cfg, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "parity": parity, "torch_gpu_baseline": tgb, "gather": gather_info,
            "kernels": kernels, "mask_decoder": f1, "alpha_clip_tower": f2, "adapter_training": f4,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
