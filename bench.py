#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: 512x512 images/sec @ 50 DDIM steps (SD-1.5 UNet, CFG 7.5, batch 4 per GPU).

  python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
  python bench.py --impl reference --gpus N --steps K ...  # the reference-side CPU path (oracle port; diffusers absent)

A "step" is one pass of the hot path over one batch: the full 50-step CFG denoising loop for 4 latents (100 UNet
sample-forwards + 50 fused CFG/DDIM updates per image) -- BASELINE config 2 per GPU; N GPUs = config 3's weak scaling
(4 images per GPU).  `value` is device-timed with latents/embeddings resident in HBM; `e2e` goes through the public
StableDiffusionPipeline.__call__ with pinned HOST buffers (H2D of embeddings + latents, D2H of the result) in the
timed region.  Data: synthetic (seeded N(0,1) latents / embeddings, random-init weights; no checkpoints offline).
One JSON line on stdout (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "512x512 images/sec @50 DDIM steps"
FLOP_PER_SAMPLE_FORWARD = 0.8033e12  # SD-1.5 @64x64 latent (SURVEY.md 8d / BASELINE.md 3)
IMAGES_PER_GPU = 4
# DRAM traffic of the GEMM family: mean of dram__bytes_read.sum + dram__bytes_write.sum over the 210 gemm2 launches of one
# batch-8 forward (one ncu pass, cold caches), see the file named here.
GEMM_DRAM_BYTES_PER_LAUNCH = 27955415   # read 27.04 MB + write 0.92 MB (outputs mostly stay in the 126 MB L2)
GEMM_DRAM_SOURCE = "profiles/r01_gemm2_dram.csv"
DDIM_STEPS = 50
GUIDANCE = 7.5


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.lines, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def fast_state_dict(expected_shapes, seed=0):
    """Random-init weights of the SD-1.5 architecture (throughput is weight-value independent); generated on the GPU."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    sd = {}
    for k, shp in expected_shapes.items():
        if ("norm" in k) and k.endswith("weight"):
            t = 1.0 + 0.05 * torch.randn(shp, generator=g, device="cuda")
        elif k.endswith("bias"):
            t = 0.02 * torch.randn(shp, generator=g, device="cuda")
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            t = torch.randn(shp, generator=g, device="cuda") * (1.0 / max(1.0, fan_in) ** 0.5)
        sd[k] = t.half()
    return sd


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner) are pointed at stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    assert args.gpus == world, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    from divergen_b200 import DDIMScheduler, StableDiffusionPipeline, UNet2DConditionModel, _lib

    unet = UNet2DConditionModel(device=dev)  # SD-1.5 config
    unet.load_state_dict(fast_state_dict(unet.expected_state_dict_shapes(), seed=0))
    sched = DDIMScheduler()
    pipe = StableDiffusionPipeline(unet, sched)

    n = IMAGES_PER_GPU
    # text embeddings: produced once on rank 0 (synthetic stand-in for CLIP output) and broadcast -- the path's only
    # data-carrying collective; seeds follow the reference's seed + rank (txt2img_...py:200).
    emb = torch.empty(2, 77, 768, dtype=torch.float16, device=dev)
    if rank == 0:
        emb.copy_(torch.randn(2, 77, 768, generator=torch.Generator().manual_seed(1234)).half())
    if world > 1:
        dist.broadcast(emb, src=0)
    pos_h = emb[0:1].repeat(n, 1, 1).cpu().pin_memory()
    neg_h = emb[1:2].repeat(n, 1, 1).cpu().pin_memory()
    gen = torch.Generator().manual_seed(42 + rank)
    lat_h = torch.randn(n, 4, 64, 64, generator=gen).half().pin_memory()
    out_h = torch.empty(n, 4, 64, 64, dtype=torch.float16).pin_memory()

    sched.set_timesteps(DDIM_STEPS)
    ts = [int(t) for t in sched.timesteps]
    al = [sched.alphas_for(t) for t in ts]
    a_t, a_prev = [a for a, _ in al], [p for _, p in al]
    ehs = torch.cat([neg_h, pos_h]).to(dev)
    lat0 = lat_h.to(dev)
    lat = lat0.clone()

    def device_step():
        lat.copy_(lat0)
        unet.denoise_loop(lat, ehs, ts, a_t, a_prev, GUIDANCE, "epsilon")

    def e2e_step():
        res = pipe(prompt_embeds=pos_h, negative_prompt_embeds=neg_h, latents=lat_h, num_inference_steps=DDIM_STEPS,
                   guidance_scale=GUIDANCE, output_type="latent").images
        out_h.copy_(res, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms = torch.tensor([e0.elapsed_time(e1), wall], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms[0].item(), ms[1].item()

    for _ in range(args.warmup):
        device_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, _ = timed(device_step, args.steps)
    launches = unet.last_launch_count * args.steps + args.steps  # + the latent reset copy per step
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    ms_e2e_dev, ms_e2e_wall = timed(e2e_step, args.steps)

    total_images = n * world * args.steps
    value = total_images / (ms_dev / 1e3)
    e2e_value = total_images / (ms_e2e_wall / 1e3)

    result = None
    if rank == 0:
        pk = peaks()
        # ---- per-kernel roofline: one eager forward with an event pair around every launch (live, this process)
        x_in = torch.cat([lat0, lat0])
        out = torch.empty_like(x_in)
        arrs = [(C.c_double * 4)() for _ in range(3)]
        launches_f = (C.c_int64 * 4)()
        total_ms = C.c_double()
        tarr = (C.c_float * 1)(float(ts[0]))
        best = None
        for _ in range(3):
            _lib.check(_lib.load().dg_unet_profile_forward(
                unet._h, C.c_void_p(x_in.data_ptr()), tarr, 1, C.c_void_p(ehs.data_ptr()), 77, C.c_void_p(out.data_ptr()),
                2 * n, 64, 64, C.c_void_p(torch.cuda.current_stream().cuda_stream), arrs[0], arrs[1], arrs[2], launches_f,
                C.byref(total_ms)), "dg_unet_profile_forward")
            cur = dict(ms=list(arrs[0]), flops=list(arrs[1]), bytes=list(arrs[2]), launches=list(launches_f), total_ms=total_ms.value)
            if best is None or cur["total_ms"] < best["total_ms"]:
                best = cur
        fam = ["gemm2_kernel(linear+conv)", "attn_tc_kernel", "groupnorm-apply", "other"]
        gemm_tf = best["flops"][0] / (best["ms"][0] * 1e-3) / 1e12
        attn_tf = best["flops"][1] / (best["ms"][1] * 1e-3) / 1e12 if best["ms"][1] else 0.0
        norm_gbs = best["bytes"][2] / (best["ms"][2] * 1e-3) / 1e9 if best["ms"][2] else 0.0
        fwd_ms_graph = ms_dev / args.steps / DDIM_STEPS
        whole_tf = FLOP_PER_SAMPLE_FORWARD * 2 * n / (fwd_ms_graph * 1e-3) / 1e12
        roofline = {
            "bound": "tensor", "kernel": "gemm2_kernel<cta_group 2, tile N 160|320> (all 210 Linear / conv1x1 / conv3x3 launches of one UNet forward)",
            "achieved": round(gemm_tf, 1), "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": round(gemm_tf / pk["tf_sustained"], 4),
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({pk['src']}; kernel timed inside a long step)",
            "traffic": GEMM_DRAM_BYTES_PER_LAUNCH, "traffic_source": GEMM_DRAM_SOURCE,
            "algorithmic_flops_per_launch": round(best["flops"][0] / max(1, best["launches"][0]) / 1e9, 2),
            "algorithmic_bytes_per_launch": round(best["bytes"][0] / max(1, best["launches"][0]) / 1e6, 2),
            "launches_per_forward": best["launches"][0], "avg_launch_us": round(best["ms"][0] * 1e3 / max(1, best["launches"][0]), 2),
            "share_of_forward": round(best["ms"][0] / best["total_ms"], 4),
            "families": {fam[i]: {"ms": round(best["ms"][i], 4), "launches": best["launches"][i],
                                  "share": round(best["ms"][i] / best["total_ms"], 4)} for i in range(4)},
            "attn_tc_tflops": round(attn_tf, 1), "attn_frac_of_peak": round(attn_tf / pk["tf_sustained"], 4),
            "norm_gbs": round(norm_gbs, 1), "norm_frac_of_hbm_peak": round(norm_gbs / pk["hbm"], 4),
            "eager_forward_ms": round(best["total_ms"], 3), "graph_forward_ms": round(fwd_ms_graph, 3),
            "whole_unet_tflops": round(whole_tf, 1), "whole_unet_frac": round(whole_tf / pk["tf_sustained"], 4),
        }
        # CPU baseline on rank 0 at N = 1 only (under torchrun OMP_NUM_THREADS=1 would cripple it; the reference arm reports it)
        cpu = cpu_baseline_sample(max_seconds=30.0) if world == 1 else None
        result = {
            "metric": METRIC, "value": round(value, 4), "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_dev / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 (fp32 accumulate / statistics / softmax)", "data": "synthetic (seeded N(0,1) latents + embeddings, random-init SD-1.5 weights)",
            "config": {"workload": "configs[1]: SD-1.5 UNet 512x512, 50 DDIM steps, CFG 7.5, batch 4 per GPU (UNet batch 8)"
                                   + (f"; x{world} GPUs = configs[2] weak scaling" if world > 1 else ""),
                       "images_per_gpu": n, "ddim_steps": DDIM_STEPS, "guidance_scale": GUIDANCE,
                       "l2": "inputs larger than L2: 1.72 GB of weights + activations stream per forward (L2 = 126 MB)",
                       "step": "one full 50-step denoising loop over the batch"},
            "e2e": {"value": round(e2e_value, 4), "unit": "images/s", "h2d_bytes_per_step": int(pos_h.nbytes + neg_h.nbytes + lat_h.nbytes),
                    "d2h_bytes_per_step": int(out_h.nbytes), "ms_per_step_wall": round(ms_e2e_wall / args.steps, 3),
                    "api": "StableDiffusionPipeline.__call__(prompt_embeds=, negative_prompt_embeds=, latents=<pinned host>, output_type='latent')"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "tensor_frac_of_roofline_images_per_s": round(value / world / (pk["tf_sustained"] * 1e12 / (FLOP_PER_SAMPLE_FORWARD * 100)), 4),
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(result), flush=True)


def _oracle_sd15():
    import torch
    from oracle.unet_oracle import UNet2DConditionOracle, UNetConfig
    torch.manual_seed(0)
    cfg = UNetConfig.sd15()
    with torch.device("meta"):
        m = UNet2DConditionOracle(cfg)
    m = m.to_empty(device="cpu").eval()
    with torch.no_grad():
        for k, p in m.named_parameters():
            if "norm" in k:
                p.fill_(1.0 if k.endswith("weight") else 0.0)
            elif p.dim() > 1:
                p.normal_(0.0, 1.0 / (p[0].numel() ** 0.5))
            else:
                p.zero_()
    return m


def cpu_baseline_sample(max_seconds=30.0, model=None, images=1, reps=None):
    """The reference-side CPU path (oracle port: diffusers itself is not installable here) on a bounded sample:
    DDIM steps of ONE image (a CFG pair through the UNet, fp32, all host threads), extrapolated x50 steps."""
    import torch
    from oracle.ddim_oracle import DDIMOracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = model or _oracle_sd15()
    s = DDIMOracle()
    s.set_timesteps(DDIM_STEPS)
    g = torch.Generator().manual_seed(42)
    lat = torch.randn(images, 4, 64, 64, generator=g)
    ehs = torch.randn(2 * images, 77, 768, generator=g)

    def one_step(i):
        nonlocal lat
        t = s.timesteps[i]
        with torch.no_grad():
            noise = m(torch.cat([lat, lat]), t, ehs).sample
        u, c = noise.chunk(2)
        lat = s.step(u + GUIDANCE * (c - u), t, lat).prev_sample

    one_step(0)  # warm-up (thread pools, allocator)
    times, i = [], 1
    t_start = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        one_step(i)
        times.append(time.perf_counter() - t0)
        i += 1
        if reps is not None:
            if len(times) >= reps:
                break
        elif time.perf_counter() - t_start > max_seconds * 0.6 or len(times) >= 5:
            break
    per_step = sum(times) / len(times)
    ips = images / (per_step * DDIM_STEPS)
    cpu_model = ""
    try:
        cpu_model = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    return {"value": round(ips, 6), "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{len(times)} DDIM step(s) of {images} image (UNet batch {2 * images}, CFG, fp32 torch CPU oracle, SD-1.5 full width), "
                      f"{per_step:.2f} s/step, extrapolated x{DDIM_STEPS} steps", "cpu_model": cpu_model,
            "note": "oracle restatement; upstream diffusers unavailable, parity unpinned"}


def run_reference(args):
    """Reference arm for this tier: the reference's CPU implementation of the path = the oracle port (diffusers is not
    vendored / pinned / installable, so there is no oracle/_ref).  Each step = one DDIM step of one image."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1; this arm is the CPU implementation and uses every host core (before torch loads)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    os.environ["MKL_NUM_THREADS"] = str(os.cpu_count() or 1)
    m = _oracle_sd15()
    t0 = time.perf_counter()
    cpu = cpu_baseline_sample(model=m, reps=max(1, args.steps))
    wall = time.perf_counter() - t0
    out = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": round(1e3 * 1.0 / (cpu["value"] * DDIM_STEPS), 3), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "configs[1]: SD-1.5 UNet 512x512, 50 DDIM steps, CFG 7.5 (bounded sample: per-DDIM-step cost of one image, extrapolated)"},
           "cpu_baseline": cpu, "e2e": {"value": cpu["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "wall_s": round(wall, 1)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
