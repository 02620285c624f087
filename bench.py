#!/usr/bin/env python
"""
bench.py -- rays/sec of the pixelNeRF render hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nproc-per-node N ... bench.py --gpus N ...   (one rank per GPU, NCCL)

A "step" renders one batch of synthetic rays (coarse + fine pass) of workload C2
(SRN-car 128x128, 2 source views, 64+32 samples, ResnetFC d=512, random-init weights with
re-randomised fc_1, real resnet34 trunk for the latent).  `value` = rays/s of the whole job
with inputs resident in HBM; `e2e` = the same through the public API
(`NeRFRenderer.bind_parallel(net)(rays)`) from pinned HOST rays to HOST pixels.  Ranks shard
rays (weak scaling: rays per GPU fixed); the scene is broadcast from rank 0 once (NCCL) and
rendered pixels are gathered to rank 0 every step inside the timed region.

`--impl reference` times the reference's algorithm on the HOST CPU cores (the oracle port of
/root/reference's PyTorch path -- the reference itself is pure Python and cannot travel to
the GPU box), on a bounded sample of the same workload.
"""
import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "pixel-nerf_b200")
sys.path.insert(0, os.path.join(PKG, "src"))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


synth = _load("pnr_synth", os.path.join(PKG, "synth.py"))

WORKLOAD = "c2"
METRIC = "rays/sec (64c+32f samples, 2 src views)"


def profiled_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture of this same command
    (profiles/r1_final_k_field_tc.txt); None if the file is absent."""
    p = os.path.join(ROOT, "profiles", "r1_final_k_field_tc.txt")
    if not os.path.exists(p):
        return None
    rd = wr = None
    n = 0
    tot = 0.0
    for line in open(p):
        if "dram__bytes_read.sum [" in line or "dram__bytes_write.sum [" in line:
            unit = line.split("[")[1].split("]")[0]
            val = float(line.split("=")[1])
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            tot += val * mult
            n += 1
    return tot / (n / 2) if n else None


def profiled_tensor_active():
    """ncu sm__pipe_tensor_cycles_active (% of peak, mean over the captured launches) from the same capture."""
    p = os.path.join(ROOT, "profiles", "r1_final_k_field_tc.txt")
    if not os.path.exists(p):
        return None
    vals = [float(l.split("=")[1]) for l in open(p) if "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed" in l]
    return sum(vals) / len(vals) if vals else None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md sustained)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = False
        self.samples = []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw,power.limit")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(float(s[0])) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(self.samples[0][1])), "reasons": reasons,
               "samples": len(sm)}
        try:    # board power next to the clocks: the kernel runs against the power limit, not the clock limit
            pw = sorted(float(s[6]) for s in self.samples if len(s) >= 8)
            if pw:
                out["power_w"] = pw[len(pw) // 2]
                out["power_limit_w"] = float(self.samples[0][7])
        except ValueError:
            pass
        return out


def model_conf(cfg):
    from util import hocon
    conf = hocon.parse_file(os.path.join(PKG, "conf", "exp", "srn.conf"))
    conf.put("model.encoder.pretrained", False)          # no network: random-init trunk
    conf.put("model.encoder.use_first_pool", cfg["use_first_pool"])
    conf.put("model.mlp_coarse.d_hidden", cfg["d_hidden"])
    conf.put("model.mlp_fine.d_hidden", cfg["d_hidden"])
    conf.put("renderer.n_coarse", cfg["n_coarse"])
    conf.put("renderer.n_fine", cfg["n_fine"])
    conf.put("renderer.n_fine_depth", cfg["n_fine_depth"])
    conf.put("renderer.white_bkgd", cfg["white_bkgd"])
    return conf


def build_scene(cfg, device, engine):
    """net (encoded, on device) + renderer through the public classes."""
    from model import make_model
    from render import NeRFRenderer
    conf = model_conf(cfg)
    torch.manual_seed(0)
    net = make_model(conf["model"])
    net.mlp_coarse.load_state_dict(synth.make_mlp_weights(11, cfg["d_hidden"]))
    net.mlp_fine.load_state_dict(synth.make_mlp_weights(12, cfg["d_hidden"]))
    net = net.to(device).eval()
    net.engine = engine
    renderer = NeRFRenderer.from_conf(conf["renderer"], eval_batch_size=50000).to(device).eval()
    src, _, focal, c = synth.make_cameras(cfg)
    images = synth.make_images(cfg, seed=0)
    with torch.no_grad():
        net.encode(images[None].to(device), src[None].to(device), focal.to(device), c=c[None].to(device))
    return net, renderer


def broadcast_scene(net, dist):
    """One NCCL broadcast of everything the render path reads (latent + cameras + weights)."""
    with torch.no_grad():
        for t in [net.encoder.latent, net.poses, net.focal, net.c, *net.mlp_coarse.parameters(),
                  *net.mlp_fine.parameters()]:
            dist.broadcast(t.data if hasattr(t, "data") else t, src=0)
        net.encoder.latent.add_(0)  # bump version -> derived state (channels-last copy) is rebuilt


def run_ours(args):
    import pnr_native as pn
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the render path has no CPU fallback)")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    cfg = synth.CONFIGS[WORKLOAD]
    net, renderer = build_scene(cfg, device, args.engine)
    if dist is not None:
        broadcast_scene(net, dist)
    render_par = renderer.bind_parallel(net, [local], simple_output=True).eval()

    n_rays = args.rays
    # every rank renders its own contiguous slice of the target orbit
    all_rays = synth.make_rays(cfg, n_rays * world, n_target=max(8, (n_rays * world) // (cfg["W"] * cfg["H"]) + 1))
    my_rays_host = all_rays[rank * n_rays:(rank + 1) * n_rays].contiguous().pin_memory()
    rays_dev = my_rays_host.to(device)[None]                      # (1, n_rays, 8) resident
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2
    gather_rgb = [torch.empty(1, n_rays, 3, device=device) for _ in range(world)] if (dist and rank == 0) else None
    gather_dep = [torch.empty(1, n_rays, device=device) for _ in range(world)] if (dist and rank == 0) else None

    def step_resident():
        flush.zero_()
        with torch.no_grad():
            rgb, depth = render_par(rays_dev)
        if dist is not None:
            dist.gather(rgb, gather_rgb, dst=0)
            dist.gather(depth, gather_dep, dst=0)
        return rgb, depth

    host_rgb = torch.empty(1, n_rays, 3).pin_memory()
    host_dep = torch.empty(1, n_rays).pin_memory()

    def step_e2e():
        flush.zero_()
        with torch.no_grad():
            r = my_rays_host.to(device, non_blocking=True)[None]
            rgb, depth = render_par(r)
        if dist is not None:
            dist.gather(rgb, gather_rgb, dst=0)
            dist.gather(depth, gather_dep, dst=0)
        host_rgb.copy_(rgb, non_blocking=True)
        host_dep.copy_(depth, non_blocking=True)

    def timed(fn, steps):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if dist is not None:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step_resident()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = pn.launch_count()
    pn.profile_begin()
    ms_total = timed(step_resident, args.steps)
    kern_ms, kern_launches = pn.profile_end()
    if os.environ.get("PNR_TC_COUNTERS"):
        names = ["mma_total", "mma_wait_a_first_chunk", "mma_wait_b", "mma_wait_bpeer", "unused4", "unused5", "mma_wait_a_later_chunks", "stream_wait_empty"]
        print("tc_counters", dict(zip(names, pn.tc_counters())), file=sys.stderr)
    launches = pn.launch_count() - launches0
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    sampler.stop_flag = True

    total_rays = n_rays * world * args.steps
    value = total_rays / (ms_total / 1e3)
    e2e_value = total_rays / (ms_e2e / 1e3)
    fl = synth.flops_per_ray(cfg["n_coarse"], cfg["n_fine"], cfg["NS"], cfg["d_hidden"])
    peak, peak_src = peaks()
    rays_per_rank_total = n_rays * args.steps
    kern_tflops = (rays_per_rank_total * fl / 1e12) / (kern_ms / 1e3) if kern_ms > 0 else None

    # fp16 tensor work the tensor engine actually issues: 3 split products over lin_in (K padded to 48) and the
    # 10 fc layers; the three lin_z GEMMs are folded into the per-scene projected-latent maps
    d = cfg["d_hidden"]
    pts = cfg["n_coarse"] + ((cfg["n_coarse"] + cfg["n_fine"]) if cfg["n_fine"] > 0 else 0)
    exec_fl = 2 * 3 * pts * (cfg["NS"] * (48 * d + 6 * d * d) + 4 * d * d)
    exec_tflops = (rays_per_rank_total * exec_fl / 1e12) / (kern_ms / 1e3) if kern_ms > 0 else None
    tensor_engine = net._fused.mlp.get("mlp_coarse", (0, 0, 0, None))[3] is not None

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_reference_run(cfg, sample_rays=args.cpu_rays, reps=1)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (fp16 hi/lo split products, fp32 accumulate)"
            if net._fused.mlp.get("mlp_coarse", (0, 0, 0, None))[3] is not None else "f32",
            "data": "synthetic",
            "config": {"workload": "C2 SRN-car 128x128, 2 src views, 64 coarse + 32 fine (16 depth) samples, "
                                   "ResnetFC d=512 x5 blocks, resnet34 latent 2x512x64x64",
                       "rays_per_step_per_gpu": n_rays, "engine": args.engine, "l2_flush_between_steps": True,
                       "parallelism": f"ray-sharded x{world}", "flop_per_ray": fl},
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": n_rays * 8 * 4 * world,
                    "d2h_bytes_per_step": n_rays * 4 * 4 * world},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": kern_tflops, "peak": peak, "unit": "TFLOP/s",
                         "frac": (kern_tflops / peak) if kern_tflops else None, "traffic": profiled_traffic(),
                         "peak_source": peak_src, "kernel_launches": int(kern_launches),
                         "ncu_tensor_pipe_active_pct": profiled_tensor_active(),
                         "traffic_note": "dram bytes per launch from profiles/r1_final_k_field_tc.txt (same command under ncu); "
                                         "dominated by the write-back / refetch caused by the 256 MB L2 flush between steps",
                         "kernel_ms_per_step": kern_ms / args.steps,
                         "executed_fp16_mma_tflops": exec_tflops if tensor_engine else None,
                         "executed_frac_of_peak": (exec_tflops / peak) if (tensor_engine and exec_tflops) else None,
                         "note": "algorithmic fp32-model FLOPs of the reference (SURVEY 8d) / device time of the "
                                 "MLP-contraction kernel(s), CUDA events on the launch stream"},
            "clocks": sampler.summary(),
        }
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def cpu_reference_run(cfg, sample_rays, reps):
    """Reference algorithm on the host cores: oracle/pnr_oracle.py (torch CPU port that mirrors
    the reference's op structure incl. its point-chunk loop), all threads."""
    oracle = _load("pnr_oracle", os.path.join(ROOT, "oracle", "pnr_oracle.py"))
    torch.set_num_threads(min(os.cpu_count(), 32))  # more threads than this only adds sync overhead here
    src, _, focal, c = synth.make_cameras(cfg)
    Hl, Wl = cfg["H"] // 2, cfg["W"] // 2
    latent = synth.make_latent(5, cfg["NS"], Hl, Wl)
    state = oracle.encode_state(src, focal, c[None], cfg["W"], cfg["H"])
    wc = synth.make_mlp_weights(11, cfg["d_hidden"])
    wf = synth.make_mlp_weights(12, cfg["d_hidden"])
    rays = synth.make_rays(cfg, sample_rays)[None]
    noise = synth.draw_noise(3, sample_rays, cfg["n_coarse"], cfg["n_fine"], cfg["n_fine_depth"])

    def once():
        with torch.no_grad():
            oracle.render(rays, noise, state, latent, wc, wf, cfg["NS"], cfg["n_coarse"], cfg["n_fine"],
                          cfg["n_fine_depth"], white_bkgd=cfg["white_bkgd"], eval_batch_size=50000)

    once()  # warm-up
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    dt = (time.perf_counter() - t0) / reps
    return {"value": sample_rays / dt, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{sample_rays} rays of the same C2 workload per pass (oracle/pnr_oracle.py, torch CPU fp32)"}


def run_torch_eager(args):
    """Context number, not the reference arm: the composed torch-op path of this repo (the same ATen op
    sequence as the reference's PyTorch code: grid_sample, 15 addmm per chunk, cat/relu/...) on ONE GPU in fp32
    (TF32 off, 50 000-point chunks) -- a stand-in for 'reference PyTorch eager on B200', which cannot be
    imported on the GPU box."""
    device = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = synth.CONFIGS[WORKLOAD]
    net, renderer = build_scene(cfg, device, "simt")
    rays = synth.make_rays(cfg, args.rays).to(device)[None]
    class TorchField:            # a generic `model` callable for NeRFRenderer's composed path
        use_viewdirs = True

        def __call__(self, xyz, coarse=True, viewdirs=None):
            return net._forward_autograd(xyz, coarse, viewdirs)

    field = TorchField()

    def step():
        with torch.no_grad():
            return renderer._forward_torch(field, rays, False)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"impl": "torch-eager", "metric": METRIC, "value": args.rays * args.steps / (ms / 1e3),
                      "unit": "rays/s", "n_gpus": 1, "steps": args.steps, "ms_per_step": ms / args.steps,
                      "dtype": "f32", "config": {"workload": "C2", "rays_per_step": args.rays,
                                                  "note": "composed torch ops of this repo's autograd path under no_grad"}}))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = synth.CONFIGS[WORKLOAD]
    oracle_stats = []
    for _ in range(max(args.warmup, 0)):
        cpu_reference_run(cfg, args.cpu_rays, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_stats.append(cpu_reference_run(cfg, args.cpu_rays, 1))
    vals = [s["value"] for s in oracle_stats]
    v = sum(vals) / len(vals)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": args.cpu_rays / v * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2 SRN-car 128x128, 2 src views, 64 coarse + 32 fine (16 depth) samples, "
                                   "ResnetFC d=512 x5 blocks", "rays_per_step": args.cpu_rays},
            "cpu_baseline": {"value": v, "unit": "rays/s", "cores": oracle_stats[0]["cores"], "kind": "port",
                             "sample": oracle_stats[0]["sample"]},
            "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-eager"])
    ap.add_argument("--engine", default=os.environ.get("PNR_ENGINE", "auto"), choices=["auto", "simt", "tc"])
    ap.add_argument("--rays", type=int, default=16384, help="rays per step per GPU (16384 = one 128x128 frame)")
    ap.add_argument("--cpu-rays", type=int, default=256, help="rays per CPU-reference pass")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "torch-eager":
        run_torch_eager(a)
    else:
        run_ours(a)
