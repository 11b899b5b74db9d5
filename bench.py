#!/usr/bin/env python
"""bench.py -- diffusion iter/s of the b200sd hot path on N B200s (one process per GPU).

A "step" is one denoising iteration of BASELINE.json configs[1] (SD-2.1-base txt2img, 512x512, 20 DDIM
steps, CFG 7.5, fp16): UNet forward at batch 2 (uncond, cond) + the fused CFG/scheduler-step kernel,
i.e. the reference's "Diffusion Speed (iter/s)" (README.md:65,83).  `value` is measured with all inputs
resident in HBM; `e2e` is the same iteration driven through the reference-facing model-call boundary
(`unet(sample=np, timestep=np, encoder_hidden_states=np)["noise_pred"]`, coreml_model.py:118-120) with
pinned HOST buffers, host<->device copies inside the timed region.  Extra keys report 512^2 images/s
(20 steps + VAE decode), the tensor-core roofline of the dominant kernel and the CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
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

import numpy as np  # noqa: E402
import torch  # noqa: E402

UNET_TFLOP = 1.609          # algorithmic FLOPs of one SD-2.1-base UNet forward at B=2
VAE_TFLOP = 2.51
PUBLISHED_ITER_S = 3.07     # best published SD-2.1-base 512^2 (iPad Pro M2, Core ML)
WORKLOAD = "SD-2.1-base txt2img 512x512, 20 DDIM steps, CFG 7.5, fp16 (BASELINE configs[1])"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d["bf16_tflops"], d["bf16_tflops_sustained"], d["hbm_gbs"], "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
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
                if v.lower().startswith("active") and not v.lower().startswith("not"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own PyTorch-CPU UNet (or its restatement when the tree is absent)
# ------------------------------------------------------------------------------------------------
def cpu_unet_runner():
    from b200sd import config
    from oracle import ref_unet, restated

    cfg = config.SD21_BASE_UNET
    sd = config.random_state_dict(config.unet_param_shapes(cfg), seed=1)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 4, 64, 64, generator=g)
    c = torch.randn(2, 1024, 1, 77, generator=g)
    t = torch.tensor([981.0, 981.0])
    torch.set_grad_enabled(False)
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm should use the host's cores (physical ~ logical / 2)
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    if ref_unet.available():
        m = ref_unet.build_unet(cfg, sd, impl="ORIGINAL")
        return "reference", (lambda: m(x, t, c)[0])
    return "port", (lambda: restated.unet_forward(sd, cfg, x, t, c))


def time_cpu(budget_s=25.0, max_steps=3, warmup=1):
    kind, fn = cpu_unet_runner()
    for _ in range(warmup):
        fn()
    times = []
    t_start = time.perf_counter()
    while len(times) < max_steps and (not times or time.perf_counter() - t_start + times[-1] < budget_s):
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
    best = min(times)
    return {"value": round(1.0 / best, 4), "unit": "iter/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{len(times)} timed fp32 UNet forwards (SD-2.1-base, bs=2, 64x64 latents, t=981; best of "
                      f"{len(times)}, median {sorted(times)[len(times) // 2]:.2f} s) after {warmup} warm-up",
            "cpu_count": os.cpu_count()}, times


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    kind, fn = cpu_unet_runner()
    t0 = time.perf_counter()
    fn()
    first = time.perf_counter() - t0
    warm = max(0, min(args.warmup, 1) - 1)  # the probe above already is one warm-up
    for _ in range(warm):
        fn()
    k = max(1, min(args.steps, int(100.0 / max(first, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(k):
        fn()
    dt = time.perf_counter() - t0
    val = k / dt
    line = {
        "impl": "reference", "metric": "diffusion_iter_per_s", "value": round(val, 4), "unit": "iter/s",
        "n_gpus": args.gpus, "steps": k, "warmup": 1 + warm, "ms_per_step": round(1e3 * dt / k, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": round(val / PUBLISHED_ITER_S, 4),
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "unet_batch": 2, "latent": "64x64", "note": "CPU arm runs the UNet forward "
                   "only (the reference does CFG + scheduler math on the host as well; negligible)"},
        "cpu_baseline": {"value": round(val, 4), "unit": "iter/s", "cores": torch.get_num_threads(), "kind": kind,
                         "sample": f"{k} fp32 UNet forwards of the reference's PyTorch-CPU path "
                                   f"({'unmodified reference modules' if kind == 'reference' else 'oracle restatement'})"
                                   f", requested steps={args.steps}"},
        "e2e": {"value": round(val, 4), "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def gemm_roofline(pipe, peaks):
    """Per-kernel roofline of the dominant kernel (umma_gemm_kernel = every conv / linear of the UNet):
    algorithmic FLOPs of each launch (2*M*N*K) / its CUDA-event duration, summed over one eager forward."""
    from b200sd import lib as L

    recs = []
    orig = L.run_gemm

    def timed(args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m = args.m if args.mode == 0 else args.n_img * (args.h // max(1, args.stride)) * (args.w // max(1, args.stride))
        k = (args.c0 + args.c1) * (9 if args.mode == 1 else 1)
        e0.record()
        orig(args)
        e1.record()
        recs.append((2.0 * m * args.n * k, e0, e1))

    unet = pipe.unet
    L.run_gemm = timed
    try:
        # park the GPU behind a ~40 ms spin kernel so the host enqueues the whole eager forward ahead of it:
        # the events then bracket back-to-back device execution instead of host launch latency
        torch.cuda.synchronize()
        torch.cuda._sleep(int(80e6))
        unet._run()  # eager (not the captured graph), same launch sequence
    finally:
        L.run_gemm = orig
    torch.cuda.synchronize()
    flops = sum(r[0] for r in recs)
    ms = sum(r[1].elapsed_time(r[2]) for r in recs)
    achieved = flops / (ms * 1e-3) / 1e12
    burst, sustained, _, how = peaks
    traffic = None  # DRAM bytes per launch of the same kernel from the committed ncu pass (profiles/traffic_r1.json)
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic_r1.json")) as f:
            traffic = round(json.load(f)["dram_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        pass
    return {"bound": "tensor", "kernel": "umma_gemm_kernel (all conv3x3 / 1x1 / linear launches of one UNet forward)",
            "achieved": round(achieved, 1), "peak": sustained, "unit": "TFLOP/s", "frac": round(achieved / sustained, 4),
            "peak_kind": f"{how} sustained bf16 dense (kernel timed inside a long step)", "traffic": traffic,
            "traffic_unit": "DRAM bytes per launch (ncu, L2 flushed per kernel; algorithmic = weights 1.73e9 B / forward)",
            "launches": len(recs), "algorithmic_tflop": round(flops / 1e12, 4), "kernel_ms_sum": round(ms, 3)}


def run_gpu_arm(args, rank, local_rank, world):
    import torch.distributed as dist
    from b200sd import lib as L
    from b200sd import scheduler as S
    from b200sd.pipeline import B200StableDiffusionPipeline

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.load()
    peaks = _peaks()
    n_steps_img, guidance = 20, 7.5
    pipe = B200StableDiffusionPipeline.from_random_init("sd21-base", images_per_call=1, device=dev, seed=1,
                                                        scheduler="DDIM")
    unet = pipe.unet
    plan = S.DDIMScheduler(n_steps_img).plan()
    g = torch.Generator().manual_seed(93 + rank)  # each rank = an independent prompt / seed
    emb_cond = torch.randn(1, 1024, 1, 77, generator=g)
    emb = torch.cat([torch.zeros_like(emb_cond), emb_cond]).half()
    lat0 = torch.randn(1, 4, 64, 64, generator=g).half().float()
    pipe._ctx.copy_(emb)
    pipe._latents.copy_(lat0)
    pipe._hist.zero_()
    coeffs = []
    for st in plan:
        k = L.StepCoeffs()
        k.guidance, k.cx, k.ce, k.x0_cx, k.x0_ce = guidance, st.cx, st.ce, st.x0_cx, st.x0_ce
        k.n_hist, k.push_eps_slot, k.push_x0_slot, k.push_x_slot = 0, -1, -1, -1
        coeffs.append((float(st.timestep), k))

    def device_step(i):
        t, k = coeffs[i % n_steps_img]
        if i % n_steps_img == 0:
            pipe._latents.copy_(lat0)
        pipe._t.fill_(t)
        sample = torch.cat([pipe._latents, pipe._latents], 0)
        npred = unet.forward_device(sample, pipe._t, pipe._ctx)
        L.cfg_scheduler_step(npred, pipe._latents, k)

    # ---- warm-up (also captures the CUDA graph) ----
    for i in range(max(3, args.warmup)):
        device_step(i)
    torch.cuda.synchronize()
    launches_per_step = (unet.launches_per_call or 0) + 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        device_step(i)
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1)
    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "iter_per_s": round(world * args.steps / (ms_dev * 1e-3), 2),
                              "ms_per_step": round(ms_dev / args.steps, 4), "launches_per_step": int(launches_per_step),
                              "pdl": os.environ.get("B200SD_PDL", "0"), "smem_kb": os.environ.get("B200SD_SMEM_KB", "")}),
                  flush=True)
        sampler.stop()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- e2e: the reference-facing boundary call with pinned host buffers, copies inside the timed region ----
    h_sample = torch.empty(2, 4, 64, 64, dtype=torch.float16).pin_memory()
    h_t = torch.empty(2, dtype=torch.float16).pin_memory()
    h_ctx = emb.clone().pin_memory()
    h_lat = lat0.clone()
    np_sample, np_t, np_ctx = h_sample.numpy(), h_t.numpy(), h_ctx.numpy()

    def e2e_step(i):
        t, k = coeffs[i % n_steps_img]
        np_sample[:] = np.concatenate([h_lat.numpy()] * 2).astype(np.float16)
        np_t[:] = t
        out = unet(sample=np_sample, timestep=np_t, encoder_hidden_states=np_ctx)["noise_pred"]  # H2D + D2H inside
        eps = out[:1] + guidance * (out[1:] - out[:1])          # host CFG + DDIM exactly like pipeline.py:559-569
        h_lat.copy_(torch.from_numpy(k.cx * h_lat.numpy() + k.ce * eps))
        if (i + 1) % n_steps_img == 0:
            h_lat.copy_(lat0)

    for i in range(3):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        e2e_step(i)
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop()

    # ---- images/s: 20 steps + VAE decode through the device-resident pipeline loop ----
    def one_image():
        final = pipe.denoise(emb, lat0, n_steps_img, guidance)
        return pipe.decode_latents(final)

    one_image()
    barrier()
    n_img = max(1, min(3, args.steps // n_steps_img + 1))
    e0.record()
    for _ in range(n_img):
        img = one_image()
    host_img = img.cpu()
    e1.record()
    barrier()
    ms_img = e0.elapsed_time(e1) / n_img

    # ---- BASELINE configs[2] shape: 8 prompts per GPU (UNet batch 16), 20 steps + VAE decode of all 8 ----
    ms_b8 = float("nan")
    if not args.no_batched:
        del one_image
        pipe8 = B200StableDiffusionPipeline.from_random_init("sd21-base", images_per_call=8, device=dev, seed=1,
                                                             scheduler="DDIM")
        emb8 = torch.cat([torch.zeros(8, 1024, 1, 77), torch.randn(8, 1024, 1, 77, generator=g)]).half()
        lat8 = torch.randn(8, 4, 64, 64, generator=g).half().float()

        def eight_images():
            return pipe8.decode_latents(pipe8.denoise(emb8, lat8, n_steps_img, guidance))

        eight_images()
        barrier()
        e0.record()
        img8 = eight_images()
        img8_host = img8.cpu()
        e1.record()
        barrier()
        ms_b8 = e0.elapsed_time(e1)

    # max over ranks
    stats = torch.tensor([ms_dev, ms_e2e, ms_img, ms_b8], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, ms_img, ms_b8 = [float(v) for v in stats.tolist()]

    if rank == 0:
        roof = gemm_roofline(pipe, peaks)
        cpu, _ = time_cpu() if world >= 1 else (None, None)
        value = world * args.steps / (ms_dev * 1e-3)
        e2e_val = world * args.steps / (ms_e2e * 1e-3)
        burst, sustained, hbm, how = peaks
        line = {
            "metric": "diffusion_iter_per_s", "value": round(value, 2), "unit": "iter/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(ms_dev / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": round(value / PUBLISHED_ITER_S, 2),
            "dtype": "f16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "unet_batch": 2, "latent": "4x64x64", "text_tokens": 77,
                       "weights": "random-init SD-2.1-base (865.9 M params)", "parallelism": f"replicas x{world}, "
                       "independent prompts per rank, no data-path collective",
                       "cache": "inputs larger than L2: every step streams 1.73 GB of fp16 weights through the 126 MB L2",
                       "cuda_graph": True},
            "e2e": {"value": round(e2e_val, 2), "unit": "iter/s", "h2d_bytes_per_step": int(np_sample.nbytes +
                    np_t.nbytes + np_ctx.nbytes), "d2h_bytes_per_step": int(2 * 4 * 64 * 64 * 4),
                    "ms_per_step": round(ms_e2e / args.steps, 4), "api": "UNetModel.__call__(**np.ndarray) boundary + "
                    "host CFG/DDIM, as in the reference loop (pipeline.py:499-573)"},
            "gpu_launches": int(launches_per_step * args.steps),
            "launches_per_step": int(launches_per_step),
            "images_per_s": round(world * 1e3 / ms_img, 3),
            "ms_per_image": round(ms_img, 2),
            "batched_images_per_s": None if ms_b8 != ms_b8 else round(world * 8 * 1e3 / ms_b8, 3),
            "batched_note": "BASELINE configs[2] shape: 8 prompts per GPU (UNet batch 16), 20 DDIM steps + VAE decode",
            "step_roofline": {"bound": "tensor", "achieved": round(UNET_TFLOP / (ms_dev / args.steps * 1e-3), 1),
                              "peak": sustained, "unit": "TFLOP/s",
                              "frac": round(UNET_TFLOP / (ms_dev / args.steps * 1e-3) / sustained, 4),
                              "note": "whole UNet forward (1.609 TFLOP algorithmic) / device time per step; per GPU"},
            "roofline": roof,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "image_checksum": float(host_img.double().sum()),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200sd", choices=["b200sd", "reference"])
    ap.add_argument("--no-batched", action="store_true", help="skip the 8-prompts-per-GPU images/s measurement")
    ap.add_argument("--quick", action="store_true", help="device-resident iter/s only (tuning runs; not a bench line)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__), "--gpus",
               str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup)]
        sys.exit(subprocess.call(cmd))
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: b200sd has no CPU fallback; use --impl reference for the CPU arm"}))
        sys.exit(1)
    run_gpu_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
