"""Benchmark of the centerOffsetRes10 hot path (BASELINE.json metric, config[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path over one batch of 64 synthetic 512x512 tiles on each GPU:
stem -> 14 tcgen05 implicit-GEMM stages -> fused heads -> decode (17 kernel launches).
Prints ONE JSON line (rank 0).  `value` = tiles/s over all GPUs with inputs resident in HBM;
`e2e` = the same through TileDetector.detect_host with pinned HOST tiles (H2D + D2H inside the timed
region); `roofline` = the dominant kernel (fused heads igemm) against the measured bf16 peak;
`cpu_baseline` = the CPU oracle (a port of the reference's PyTorch path) on this box's host cores.
`--impl reference` times that CPU path alone.
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

METRIC = "512x512 tiles/sec centerOffsetRes10 infer+decode"
FLOPS_PER_TILE = 49.2957e9               # SURVEY.md 8a (18 convs + 3 deconvs, deconv without zero insertion)
HEADS_FLOPS_PER_TILE = 2.0 * 16384 * (384 * 2304 + 7 * 128)
WORKLOAD = "configs[1]: centerOffsetRes10 batched inference+decode, batch 64 of 512x512 tiles per GPU, bf16"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons.  The sampler runs from before the warm-up to after the timed
    region (nvidia-smi needs a few hundred ms to start); samples are matched to the timed window by time."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        ok = [(t, r) for t, r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        win = [r for t, r in ok if t0 - 0.02 <= t <= t1 + 0.05]
        window = "timed region"
        if not win:                       # timed region shorter than the sampling period
            win, window = [r for t, r in ok if t >= t0 - 2.0], "warm-up + timed region"
        sm = [float(r[0]) for r in win]
        mx = [float(r[1]) for r in win]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i] == "Active" for r in win)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "window": window}


def cpu_oracle_rate(tiles_per_iter, min_seconds, warmup=1, max_iters=1000, fixed_iters=None):
    """tiles/s of the CPU oracle (reference port: ATen fp32 on host cores) for infer + decode."""
    import torch
    from oracle import centernet_cpu as O
    torch.set_num_threads(os.cpu_count() or 1)
    sd = O.make_state_dict(1234)
    x = O.make_tiles(tiles_per_iter, seed=0)

    def step():
        with torch.no_grad():
            out = O.resnet10_forward(sd, x)[0]
            O.decode_centernet(out, K=100)

    for _ in range(warmup):
        step()
    times = []
    t0 = time.perf_counter()
    while True:
        t1 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t1)
        if fixed_iters is not None:
            if len(times) >= fixed_iters:
                break
        elif time.perf_counter() - t0 >= min_seconds or len(times) >= max_iters:
            break
    total = sum(times)
    return tiles_per_iter * len(times) / total, total / len(times), len(times), torch.get_num_threads()


def run_reference(args, rank):
    """The reference arm: the reference's own CPU implementation of the path.  The reference is pure
    Python/PyTorch (nothing to compile), and /root/reference does not exist on the GPU box, so this runs
    the oracle port (oracle/centernet_cpu.py, pinned to the reference by tests/golden) on all host cores."""
    if rank != 0:
        return
    sample = 8
    rate, sec_per_iter, iters, cores = cpu_oracle_rate(sample, 0, warmup=max(1, min(args.warmup, 2)),
                                                       fixed_iters=args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "tiles/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_iter * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": "%d tiles per step on the host CPU" % sample},
            "cpu_baseline": {"value": rate, "unit": "tiles/s", "cores": cores, "kind": "port",
                             "sample": "%d steps x %d tiles, infer+decode, fp32 ATen on %d threads"
                                       % (iters, sample, cores)},
            "e2e": {"value": rate, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="tiles per GPU per step (config[1] = 64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--slide", type=int, default=16384, help="edge of the synthetic whole-slide image (configs[4]); 0 = skip")
    ap.add_argument("--hbm-kernels", type=int, default=2048, help="tiles for the decode / loss / render roofline leg; 0 = skip")
    ap.add_argument("--train-steps", type=int, default=10, help="timed training steps (configs[2]/[3]); 0 = skip")
    ap.add_argument("--train-batch", type=int, default=32, help="samples per GPU per training step (exp.json)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    import scd_resnet_b200 as S           # raises if libscd_b200.so is missing: no fallback
    from scd_resnet_b200 import synthetic
    from scd_resnet_b200.centerNetOffset import CenterNetResidual
    from scd_resnet_b200.inference import TileDetector

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B, K, W = args.batch, args.steps, args.warmup
    model = CenterNetResidual(10)
    model.load_state_dict(synthetic.make_state_dict(model, 1234))
    model.eval()
    det = TileDetector(model, B, dev)
    # three rotating input batches (3 x 64 MB) and 1.4 GB of activations: far beyond the 126 MB L2
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    xs = [torch.randn(B, 1, 512, 512, device=dev, generator=g) for _ in range(3)]

    # ---- device-resident throughput --------------------------------------------------------------
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for i in range(W):
        det.detect_device(xs[i % 3])
    stage_ev = [[torch.cuda.Event(enable_timing=True) for _ in range(17)] for _ in range(K)]
    dec_ev = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    for evs in stage_ev:
        for e in evs:
            e.record()                    # materialise the cudaEvent_t handles
    for e in dec_ev:
        e.record()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    t_start.record()
    for i in range(K):
        det.detect_device(xs[i % 3], stage_ev[i])
        dec_ev[i].record()
    t_end.record()
    barrier()
    ms = t_start.elapsed_time(t_end)
    clk = clocks.stop(wall0, time.time()) if rank == 0 else None
    stage_ms = [statistics.mean(stage_ev[i][j].elapsed_time(stage_ev[i][j + 1]) for i in range(K)) for j in range(16)]
    decode_ms = statistics.mean(stage_ev[i][16].elapsed_time(dec_ev[i]) for i in range(K))

    # ---- the same step with fp16 operands (same tensor-core rate, 8x finer mantissa: the mode that meets the
    # 1e-2 parity bar on every head; profiles/accuracy_*.json) -------------------------------------------
    det16 = TileDetector(model, B, dev, precision="fp16")
    for i in range(W):
        det16.detect_device(xs[i % 3])
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    f0.record()
    for i in range(K):
        det16.detect_device(xs[i % 3])
    f1.record()
    barrier()
    ms16 = f0.elapsed_time(f1)
    del det16

    # ---- end to end: pinned host tiles in, host detections out ---------------------------------------
    host = [torch.randn(B, 1, 512, 512).pin_memory() for _ in range(3)]
    det.detect_host([host[i % 3] for i in range(3)])
    barrier()
    t0 = time.perf_counter()
    out = det.detect_host([host[i % 3] for i in range(K)])
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    # device-side bracket: first upload enqueued -> last download finished (events on the copy streams); the
    # host wall clock around the same call is reported too and the slower of the two is used
    e2e_dev_ms = det.t_first.elapsed_time(det.t_last)
    barrier()
    e2e_ms = max(e2e_dev_ms, e2e_wall_ms)

    # ---- whole slide (configs[4]): 16384 x 16384 synthetic slide (uint8, pinned host) -> upload -> on-device
    # reflect pad / stride-384 tiling / fp64 normalise -> 1849 tiles sharded over the ranks -> decode -> all-gather
    # of the per-tile detections -> ordered merge on the host (test.py:41-142) -----------------------------
    slide_info = None
    if args.slide > 0:
        from scd_resnet_b200 import slide as slide_mod
        gs = torch.Generator().manual_seed(7)
        gray = torch.randint(0, 256, (args.slide, args.slide), dtype=torch.uint8, generator=gs).pin_memory()
        slide_mod.analyse_slide(det, gray, group=dist.group.WORLD if world > 1 else None)        # warm-up
        barrier()
        t0 = time.perf_counter()
        dets, planes = slide_mod.analyse_slide(det, gray, group=dist.group.WORLD if world > 1 else None)
        barrier()
        slide_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([slide_s], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            slide_s = float(t.item())
        slide_info = {"workload": "configs[4]: %d x %d slide, %d overlapping 512x512 tiles, global merge"
                                  % (args.slide, args.slide, planes.shape[1]),
                      "seconds": slide_s, "tiles_per_s": planes.shape[1] / slide_s, "tiles": int(planes.shape[1]),
                      "detections": int(dets.shape[0]), "h2d_bytes": int(gray.numel()),
                      "timing": "host wall clock around analyse_slide (upload, tiling, inference, decode, gather, merge), max over ranks"}
        del gray, planes

    # ---- the HBM-bound kernels of the path at a size where a roofline fraction means something (2048 tiles /
    # samples; at the batch sizes of configs[1..3] they move 4-8 MB and are launch bound).  L2 flushed between
    # launches; algorithmic bytes per unit from SURVEY.md 8d. --------------------------------------------
    hbm_kernels = None
    if rank == 0 and args.hbm_kernels > 0:
        N = args.hbm_kernels
        peaks_ = measured_peaks()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def timeit(fn, iters=15):
            for _ in range(3):
                fn()
            ts = []
            for _ in range(iters):
                flush.zero_()
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b_.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b_))
            return statistics.median(ts)

        hh = torch.randn(N, 1, 128, 128, device=dev, generator=g) * 1.5 - 2
        rr = torch.randn(N, 4, 128, 128, device=dev, generator=g)
        oo = torch.randn(N, 2, 128, 128, device=dev, generator=g)
        ll, cc = synthetic.make_objects(N, seed=1)
        ll, cc = ll.to(dev), cc.to(dev)
        gt = S.ops.render_targets(ll, cc, with_npos=True)
        res = {}
        for name, unit_bytes, fn in (
                ("decode", 73136, lambda: S.ops.decode_topk(hh, rr, oo, K=100)),
                ("loss_fwd_bwd", 3 * 65536 + 30 * 6 * 4 * 2, lambda: S.ops.centernet_loss_sparse(hh, rr, oo, *gt[:4], npos=gt[4])),
                ("render_targets", 65536 + 960, lambda: S.ops.render_targets(ll, cc, with_npos=True))):
            t_ms = timeit(fn)
            gbs = unit_bytes * N / t_ms / 1e6
            res[name] = {"units": N, "ms": round(t_ms, 4), "alg_bytes_per_unit": unit_bytes, "achieved_gbs": round(gbs, 1),
                         "frac_of_hbm_peak": round(gbs / peaks_["hbm_gbs"], 4)}
        # data path (row f3): 256 samples of a 512-tile resident dataset; 1 MB tile + 1 MB noise read, 1 MB written
        NA = 256
        ds = torch.randint(0, 256, (512, 512, 512), device=dev, generator=g).float()
        dl, dc = synthetic.make_objects(512, seed=2)
        dl, dc = dl.to(dev), dc.to(dev)
        ai = torch.randperm(512, device=dev, generator=g)[:NA]
        af = torch.rand(NA, 2, device=dev, generator=g) > 0.5
        aj = torch.randn(NA, device=dev, generator=g)
        an = torch.randn(NA, 512, 512, device=dev, generator=g)
        t_ms = timeit(lambda: S.ops.augment_batch(ds, dl, dc, ai, af, aj, an))
        gbs = 3 * (1 << 20) * NA / t_ms / 1e6
        res["augment_batch"] = {"units": NA, "ms": round(t_ms, 4), "alg_bytes_per_unit": 3 << 20, "achieved_gbs": round(gbs, 1),
                                "frac_of_hbm_peak": round(gbs / peaks_["hbm_gbs"], 4)}
        del ds, an
        hbm_kernels = {"peak_gbs": peaks_["hbm_gbs"], "peak_source": peaks_["source"], "l2": "256 MB flush before every launch",
                       "timing": "CUDA events around the public op (all its launches and memsets)", **res}
        del hh, rr, oo, gt, flush

    # ---- training step (configs[2] / [3]): render targets + forward + loss + backward + Adam ---------------
    train_ms = None
    if args.train_steps > 0:
        from scd_resnet_b200.training import TrainEngine
        del det, xs
        torch.cuda.empty_cache()
        TB = args.train_batch
        tmodel = CenterNetResidual(10)
        tmodel.load_state_dict(synthetic.make_state_dict(tmodel, 1234))
        tmodel.to(dev).train()
        eng = TrainEngine(tmodel, process_group=dist.group.WORLD if world > 1 else None,
                          peer_stats=os.environ.get("SCD_PEER_STATS", "1") != "0")
        txs = [torch.randn(TB, 1, 512, 512, device=dev, generator=g) for _ in range(3)]
        tlocs = [tuple(t.to(dev) for t in synthetic.make_objects(TB, seed=50 + 3 * rank + i)) for i in range(3)]

        def train_once(i):
            ys = S.ops.render_targets(*tlocs[i % 3], with_npos=True)
            return eng.train_step(txs[i % 3], ys)

        for i in range(3):
            train_once(i)
        barrier()
        ts, te = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts.record()
        for i in range(args.train_steps):
            last = train_once(i)
        te.record()
        barrier()
        train_ms = ts.elapsed_time(te)
        train_loss = float(last[0])

    if world > 1:
        t = torch.tensor([ms, e2e_ms, train_ms or 0.0, ms16], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms, tm, ms16 = t.tolist()
        train_ms = tm if train_ms is not None else None

    if rank == 0:
        peaks = measured_peaks()
        heads_ms = stage_ms[15]
        ach = HEADS_FLOPS_PER_TILE * B / (heads_ms * 1e-3) / 1e12
        names = ["stem", "l1c1", "l1c2", "l2ds", "l2c1", "l2c2", "l3ds", "l3c1", "l3c2", "l4ds", "l4c1", "l4c2",
                 "dc1", "dc2", "dc3", "heads"]
        line = {
            "metric": METRIC, "value": world * B * K / (ms * 1e-3), "unit": "tiles/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "tile": 512, "K": 100,
                       "weights": "synthetic He-init, BN folded (seed 1234)",
                       "l2": "3 rotating input batches (3x64 MB) + 1.4 GB activations per step > 126 MB L2"},
            "e2e": {"value": world * B * K / (e2e_ms * 1e-3), "unit": "tiles/s",
                    "h2d_bytes_per_step": B * 512 * 512 * 4, "d2h_bytes_per_step": 10 * B * 100 * 4,
                    "api": "TileDetector.detect_host (pinned host tiles -> host detections, copies overlapped)",
                    "h2d_gb_per_s": round(B * 512 * 512 * 4 / (e2e_ms / K * 1e-3) / 1e9, 2),
                    "note": "fp32 tiles are 1 MB each: when this falls below `value` the host -> device link of the box is "
                            "the limit (the copies overlap the kernels), not the GPU"},
            "gpu_launches": 17 * K,
            "roofline": {"kernel": "igemm_kernel<384, EPI_HEADS> (fused heads)", "bound": "tensor",
                         "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": ach / peaks["bf16_tflops_sustained"],
                         "traffic": 569620992 * B // 64,    # dram read + write per launch, ncu --set full (profiles/ncu_full_r01_c.json)
                         "peak_source": peaks["source"] + " sustained bf16 (kernel timed inside a long step)",
                         "ms_per_launch": heads_ms},
            "step_tflops": FLOPS_PER_TILE * B / (ms / K * 1e-3) / 1e12,
            "fp16_operands": {"value": world * B * K / (ms16 * 1e-3), "unit": "tiles/s", "ms_per_step": ms16 / K,
                              "note": "same step with precision='fp16' (rel-RMS vs fp32 ~1e-3; bf16 0.6-1.3e-2)"},
            "stage_ms": {n: round(v, 4) for n, v in zip(names, stage_ms)}, "decode_ms": round(decode_ms, 4),
            "clocks": clk,
        }
        if slide_info is not None:
            line["slide"] = slide_info
        if hbm_kernels is not None:
            line["hbm_kernels"] = hbm_kernels
        if train_ms is not None:
            sps = world * args.train_batch * args.train_steps / (train_ms * 1e-3)
            line["train"] = {"metric": "training samples/sec, centerOffsetRes10, batch 32 per GPU (exp.json)",
                             "value": sps, "unit": "samples/s", "ms_per_step": train_ms / args.train_steps,
                             "steps": args.train_steps, "batch_per_gpu": args.train_batch,
                             "step": "render targets + fwd (batch-stat BN) + focal/L1 loss + bwd + Adam",
                             "grad_sync": ("NCCL all-reduce of the flat fp32 gradient in 3 ranges started during the backward pass; BN statistics: "
                                           + ("one-shot NVLink peer-memory all-reduce kernel" if eng.peer is not None
                                              else "NCCL all-reduce (%s)" % eng.peer_reason)) if world > 1 else "none",
                             "tflops": 147.5e9 * args.train_batch / (train_ms / args.train_steps * 1e-3) / 1e12,
                             "last_loss": train_loss}
        if world == 1 and not args.no_cpu_baseline:
            rate, spi, iters, cores = cpu_oracle_rate(8, 10.0)
            line["cpu_baseline"] = {"value": rate, "unit": "tiles/s", "cores": cores, "kind": "port",
                                    "sample": "%d iterations x 8 tiles, infer+decode, fp32 ATen on %d threads"
                                              % (iters, cores)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
