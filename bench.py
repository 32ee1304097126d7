"""Benchmark of the centerOffsetRes10 hot path (BASELINE.json metric, config[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path over one batch of 64 synthetic 512x512 tiles on each GPU:
stem -> 14 tcgen05 implicit-GEMM stages -> fused heads -> decode (17 kernel launches).
Prints ONE JSON line (rank 0).  `value` = tiles/s over all GPUs with inputs resident in HBM;
`e2e` = the same through TileDetector.detect_host with pinned HOST tiles (H2D + D2H inside the timed
region), timed for both input forms (grey bytes / float32 tiles), the faster one on this box as the headline; `roofline` = the dominant kernel (fused heads igemm) against the measured bf16 peak;
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
# arithmetic of the timed path: the bf16 model (every weight rounded to bf16) x fp16 activations, fp32 accumulation on
# tcgen05 kind::f16 (weights.PRECISIONS["mixed"]: all three heads within the north star's 1e-2 of the fp32 reference;
# the pure-bf16 and pure-fp16 plans are timed next to it)
DTYPE = "bf16 weights x fp16 activations, fp32 accumulate"


def e2e_entry(world, B, K, u8_ms, f32_ms):
    """The end-to-end line: both input forms of TileDetector.detect_host are timed (grey bytes normalised on the device, a
    quarter of the bytes over the host link; float32 tiles normalised on the host, what the reference's test.py uploads);
    the headline is the faster one on this box, the other is reported beside it."""
    forms = {
        "grey_bytes": {"value": world * B * K / (u8_ms * 1e-3), "h2d_bytes_per_step": B * 512 * 512,
                       "h2d_gb_per_s": round(B * 512 * 512 / (u8_ms / K * 1e-3) / 1e9, 2),
                       "note": "uint8 tiles as the slide holds them; the per-tile fp64 normalisation (test.py:89) runs on the "
                               "device (scd_tiles_normalize_u8, +0.05 ms per 64 tiles)"},
        "float32_tiles": {"value": world * B * K / (f32_ms * 1e-3), "h2d_bytes_per_step": B * 512 * 512 * 4,
                          "h2d_gb_per_s": round(B * 512 * 512 * 4 / (f32_ms / K * 1e-3) / 1e9, 2),
                          "note": "tiles normalised on the host arrive as float32, 1 MB each (the reference's own form): needs "
                                  ">= 24 GB/s of host link per GPU to stay compute bound"},
    }
    best = max(forms, key=lambda k: forms[k]["value"])
    return {"value": forms[best]["value"], "unit": "tiles/s", "h2d_bytes_per_step": forms[best]["h2d_bytes_per_step"],
            "d2h_bytes_per_step": 10 * B * 100 * 4, "form": best,
            "api": "TileDetector.detect_host (pinned host tiles -> [per-tile normalise on the device] -> inference + decode -> "
                   "host detections; copies overlapped with the kernels on three streams)",
            "grey_bytes": forms["grey_bytes"], "float32_tiles": forms["float32_tiles"]}


def workload_config(batch):
    return {"workload": WORKLOAD, "batch_per_gpu": batch, "tile": 512, "K": 100,
            "weights": "synthetic He-init, BN folded (seed 1234)",
            "l2": "3 rotating input batches (3x64 MB) + 1.4 GB activations per step > 126 MB L2"}


def heads_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the heads kernel per launch at batch 64, from the committed
    `ncu --set full` capture of this round (profiles/ncu_full_r02.json), else the round-1 capture; None if neither."""
    for name in ("ncu_full_r02.json", "ncu_full_r01_c.json"):
        p = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(p):
            continue
        try:
            with open(p) as f:
                rows = json.load(f)
            rows = rows["kernels"] if isinstance(rows, dict) and "kernels" in rows else rows

            def nbytes(v):              # tools/ncu_summary.py keeps ncu's "<value> <unit>" strings
                num, unit = (str(v).split() + ["byte"])[:2]
                return float(num) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]

            for r in rows:
                kn = r.get("kernel", r.get("name", ""))
                if "igemm_kernel<384" in kn or "igemm_kernel<(int)384" in kn:
                    rd, wr = r.get("dram__bytes_read.sum"), r.get("dram__bytes_write.sum")
                    if rd is not None and wr is not None:
                        return int(nbytes(rd) + nbytes(wr)), "profiles/" + name
        except Exception:
            continue
    return None, None


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


def cpu_oracle_rate(tiles_per_iter, min_seconds, warmup=1, max_iters=1000, fixed_iters=None, chunk=8):
    """tiles/s of the CPU oracle (reference port: ATen fp32 on host cores) for infer + decode.  One iteration =
    tiles_per_iter tiles, run in chunks of `chunk` (the reference's own test.py batches 24; memory stays bounded)."""
    import torch
    from oracle import centernet_cpu as O
    torch.set_num_threads(os.cpu_count() or 1)
    sd = O.make_state_dict(1234)
    xs = [O.make_tiles(min(chunk, tiles_per_iter - c), seed=c) for c in range(0, tiles_per_iter, chunk)]

    def step():
        with torch.no_grad():
            for x in xs:
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
    # a step = the workload's batch (64 tiles) as long as the whole run stays within a few minutes on ~20 tiles/s of
    # host throughput; more than 30 steps shrink the per-step sample (said in cpu_baseline.sample)
    sample = args.batch if args.steps <= 30 else max(8, (args.batch * 30 // args.steps) // 8 * 8)
    rate, sec_per_iter, iters, cores = cpu_oracle_rate(sample, 0, warmup=max(1, min(args.warmup, 2)),
                                                       fixed_iters=args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "tiles/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_iter * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.batch),
            "cpu_baseline": {"value": rate, "unit": "tiles/s", "cores": cores, "kind": "port",
                             "sample": "%d steps x %d tiles, infer+decode, fp32 ATen on %d threads"
                                       % (iters, sample, cores)},
            "e2e": {"value": rate, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400, help="timed steps (400 x 2.7 ms > 1 s of timed region)")
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="tiles per GPU per step (config[1] = 64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the cuDNN yardstick leg (oracle graph on CUDA)")
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
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    t_start.record()
    for i in range(K):
        det.detect_device(xs[i % 3])          # no events inside the step: the kernels chain by programmatic dependent launch
    t_end.record()
    barrier()
    ms = t_start.elapsed_time(t_end)
    clk = clocks.stop(wall0, time.time()) if rank == 0 else None
    # per-stage times from a separate pass with an event between every two launches (which serialises the launches: the
    # stage times add up to slightly more than ms_per_step)
    KS = min(K, 30)
    stage_ev = [[torch.cuda.Event(enable_timing=True) for _ in range(17)] for _ in range(KS)]
    dec_ev = [torch.cuda.Event(enable_timing=True) for _ in range(KS)]
    for evs in stage_ev:
        for e in evs:
            e.record()                    # materialise the cudaEvent_t handles
    for e in dec_ev:
        e.record()
    for i in range(KS):
        det.detect_device(xs[i % 3], stage_ev[i])
        dec_ev[i].record()
    torch.cuda.synchronize()
    stage_ms = [statistics.mean(stage_ev[i][j].elapsed_time(stage_ev[i][j + 1]) for i in range(KS)) for j in range(16)]
    decode_ms = statistics.mean(stage_ev[i][16].elapsed_time(dec_ev[i]) for i in range(KS))

    # ---- the same step under the other two precision plans (same kernels, same tensor-core rate): pure bf16 (weights
    # AND activations; offset head 1.26e-2 off the fp32 reference) and pure fp16 (1e-3) ------------------------
    K2 = min(K, 50)
    other_ms = {}
    for plan in ("bf16", "fp16"):
        det2 = TileDetector(model, B, dev, precision=plan)
        for i in range(W):
            det2.detect_device(xs[i % 3])
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        f0.record()
        for i in range(K2):
            det2.detect_device(xs[i % 3])
        f1.record()
        barrier()
        other_ms[plan] = f0.elapsed_time(f1) / K2
        del det2

    # ---- end to end: pinned host tiles in, host detections out ---------------------------------------
    # Tiles arrive the way the reference's pipeline holds them before normalize (test.py:21-33, 89): grey BYTES; the
    # per-tile fp64 normalisation runs on the device (scd_tiles_normalize_u8).  The float32 form (tiles normalised on
    # the host, 4x the bytes over the host link) is timed as well.
    def e2e_run(host):
        det.detect_host([host[i % 3] for i in range(3)])
        best = None
        for _ in range(2):                      # two passes of K steps, the better one: a host hiccup (page faults of the
            barrier()                           # freshly pinned buffers, another rank's setup) is not the pipeline's rate
            t0 = time.perf_counter()
            det.detect_host([host[i % 3] for i in range(K)])
            wall_ms = (time.perf_counter() - t0) * 1e3
            # device-side bracket: first upload enqueued -> last download finished (events on the copy streams); the
            # host wall clock around the same call is taken too and the slower of the two is used
            dev_ms = det.t_first.elapsed_time(det.t_last)
            barrier()
            ms = max(dev_ms, wall_ms)
            best = ms if best is None else min(best, ms)
        return best

    gh = torch.Generator().manual_seed(11 + rank)
    host_u8 = [torch.randint(0, 256, (B, 1, 512, 512), dtype=torch.uint8, generator=gh).pin_memory() for _ in range(3)]
    e2e_ms = e2e_run(host_u8)
    del host_u8
    host_f32 = [torch.randn(B, 1, 512, 512, generator=gh).pin_memory() for _ in range(3)]
    e2e_f32_ms = e2e_run(host_f32)
    del host_f32

    # ---- whole slide (configs[4]): 16384 x 16384 synthetic slide (uint8, pinned host) -> upload -> on-device
    # reflect pad / stride-384 tiling / fp64 normalise -> 1849 tiles sharded over the ranks -> decode -> all-gather
    # of the per-tile detections -> ordered merge on the host (test.py:41-142) -----------------------------
    slide_info = None
    if args.slide > 0:
        from scd_resnet_b200 import slide as slide_mod
        gs = torch.Generator().manual_seed(7)
        gray = torch.randint(0, 256, (args.slide, args.slide), dtype=torch.uint8, generator=gs).pin_memory()
        grp = dist.group.WORLD if world > 1 else None
        slide_mod.analyse_slide(det, gray, group=grp, return_planes=False)        # warm-up
        barrier()
        t0 = time.perf_counter()
        dets, _ = slide_mod.analyse_slide(det, gray, group=grp, return_planes=False)
        barrier()
        slide_s = time.perf_counter() - t0
        n_slide_tiles = S.ops.slide_geometry(args.slide, args.slide)[0] * S.ops.slide_geometry(args.slide, args.slide)[1]
        if world > 1:
            t = torch.tensor([slide_s], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            slide_s = float(t.item())
        slide_info = {"workload": "configs[4]: %d x %d slide, %d overlapping 512x512 tiles, global merge"
                                  % (args.slide, args.slide, n_slide_tiles),
                      "seconds": slide_s, "tiles_per_s": n_slide_tiles / slide_s, "tiles": int(n_slide_tiles),
                      "detections": int(dets.shape[0]),
                      "h2d_bytes": "each rank uploads only the column strip its tile columns read (%d bytes for the whole slide)" % int(gray.numel()),
                      "flow": "pinned uint8 slide -> per-tile-column strip upload overlapped with compute -> on-device reflect pad / "
                              "tiling / fp64 normalise -> inference + decode -> on-device threshold + coordinate merge -> all-gather of "
                              "the kept rows",
                      "timing": "host wall clock around analyse_slide, max over ranks"}
        del gray

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
        train_extra = {}
        if world > 1:
            # N-rank training follows the single-rank run (DDP + SyncBatchNorm semantics, ref: networkFactory.py:133-134):
            # 3 steps on IDENTICAL data on every rank (the global batch is the local batch repeated: same statistics, same
            # mean gradient) from the same initial weights, once through the N-rank engine and once through a local one
            def three_steps(group):
                m = CenterNetResidual(10)
                m.load_state_dict(synthetic.make_state_dict(m, 1234))
                m.to(dev).train()
                e = TrainEngine(m, process_group=group, peer_stats=os.environ.get("SCD_PEER_STATS", "1") != "0")
                gx = torch.Generator(device=dev).manual_seed(4242)
                cx = torch.randn(8, 1, 512, 512, device=dev, generator=gx)
                cl = tuple(t.to(dev) for t in synthetic.make_objects(8, seed=77))
                p0 = e.P.clone()
                ls = [e.train_step(cx, S.ops.render_targets(*cl, with_npos=True)).clone() for _ in range(3)]
                torch.cuda.synchronize()
                return torch.stack(ls).cpu(), (e.P - p0).double()

            def cosine(a, b):
                return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))

            ln, un = three_steps(dist.group.WORLD)
            l1, u1 = three_steps(None)
            l2, u2 = three_steps(None)          # the yardstick: two single-rank runs differ by the order of their atomics
            rel = lambda a, b: float(((a - b).abs() / b.abs().clamp_min(1e-12)).max())
            t = torch.tensor([rel(ln, l1), rel(l2, l1), -cosine(un, u1), -cosine(u2, u1)], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dl_n, dl_s, cos_n, cos_s = float(t[0]), float(t[1]), -float(t[2]), -float(t[3])
            train_extra["ddp_check"] = {
                "what": "3 Adam steps on identical data from the same weights: N-rank engine vs a single-rank engine on every "
                        "rank (worst rank), next to single-rank vs single-rank (run-to-run noise of the fp32 / fp64 atomics)",
                "max_rel_loss_diff": dl_n, "max_rel_loss_diff_single_vs_single": dl_s,
                "update_cosine": cos_n, "update_cosine_single_vs_single": cos_s,
                "ok": bool(dl_n <= max(5e-3, 3.0 * dl_s) and cos_n >= cos_s - 0.03)}
        elif rank == 0:
            # the drop-in route: the reference's loop body (zero_grad -> model() -> loss -> backward -> torch Adam step,
            # ref: networkFactory.py:257-263) on the plugin module; same kernels through one autograd node
            from scd_resnet_b200.centerNetOffset import CenterNetLoss
            amodel = CenterNetResidual(10)
            amodel.load_state_dict(synthetic.make_state_dict(amodel, 1234))
            opt = torch.optim.Adam(filter(lambda p: p.requires_grad, amodel.parameters()))
            amodel.to(dev).train()
            lossfn = CenterNetLoss(0.1, 0.1)

            def auto_once(i):
                ys = S.ops.render_targets(*tlocs[i % 3], with_npos=True)
                opt.zero_grad()
                loss, _ = lossfn(amodel(txs[i % 3], decode=False), list(ys))
                loss = loss.mean()
                loss.backward()
                opt.step()
                return loss

            for i in range(3):
                auto_once(i)
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for i in range(args.train_steps):
                auto_once(i)
            a1.record()
            torch.cuda.synchronize()
            ams = a0.elapsed_time(a1) / args.train_steps
            train_extra["autograd_route"] = {"ms_per_step": ams, "samples_per_s": args.train_batch / (ams * 1e-3),
                                             "what": "reference loop body verbatim on the plugin module + torch.optim.Adam "
                                                     "(foreach) on the parameter views"}
            del amodel, opt

    if world > 1:
        t = torch.tensor([ms, e2e_ms, train_ms or 0.0, other_ms["bf16"], other_ms["fp16"], e2e_f32_ms], device=dev,
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms, tm, other_ms["bf16"], other_ms["fp16"], e2e_f32_ms = t.tolist()
        train_ms = tm if train_ms is not None else None

    # ---- the reference's own GPU path on this box: the oracle (the reference's ATen graph) on CUDA through cuDNN, fp32
    # and bf16 autocast + channels_last, same batch, forward + decode; a yardstick next to `value`, never the product
    gpu_reference = None
    if rank == 0 and world == 1 and not args.no_gpu_reference:
        from oracle import centernet_cpu as O
        try:
            torch.cuda.empty_cache()
            sdg = {k: v.to(dev) for k, v in synthetic.make_state_dict(CenterNetResidual(10), 1234).items()}
            xr = torch.randn(B, 1, 512, 512, device=dev)
            gpu_reference = {"what": "oracle forward + decode on CUDA (cuDNN / ATen), batch %d, eval mode" % B}
            for name, ctx, cl in (("fp32", None, False), ("bf16_autocast_channels_last", torch.bfloat16, True)):
                sdx = {k: (v.contiguous(memory_format=torch.channels_last) if (cl and v.dim() == 4) else v) for k, v in sdg.items()}
                xin = xr.contiguous(memory_format=torch.channels_last) if cl else xr

                def ref_step():
                    with torch.no_grad():
                        if ctx is None:
                            out = O.resnet10_forward(sdx, xin)[0]
                        else:
                            with torch.autocast("cuda", dtype=ctx):
                                out = O.resnet10_forward(sdx, xin)[0]
                        O.decode_centernet({k: v.float() for k, v in out.items()}, K=100)

                torch.backends.cudnn.benchmark = True
                for _ in range(3):
                    ref_step()
                torch.cuda.synchronize()
                r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n_it = 10
                r0.record()
                for _ in range(n_it):
                    ref_step()
                r1.record()
                torch.cuda.synchronize()
                rms = r0.elapsed_time(r1) / n_it
                gpu_reference[name] = {"ms_per_step": rms, "tiles_per_s": B / (rms * 1e-3)}
            del sdg, xr
        except Exception as e:                                 # a yardstick must not take the bench line down
            gpu_reference = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank == 0:
        peaks = measured_peaks()
        heads_ms = stage_ms[15]
        ach = HEADS_FLOPS_PER_TILE * B / (heads_ms * 1e-3) / 1e12
        names = ["stem", "l1c1", "l1c2", "l2ds", "l2c1", "l2c2", "l3ds", "l3c1", "l3c2", "l4ds", "l4c1", "l4c2",
                 "dc1", "dc2", "dc3", "heads"]
        # a kernel timed inside a region of >= 1 s sees the sustained clocks, a shorter region the burst clocks
        long_region = ms >= 1000.0
        peak = peaks["bf16_tflops_sustained"] if long_region else peaks["bf16_tflops"]
        traffic, traffic_src = heads_traffic()
        line = {
            "metric": METRIC, "value": world * B * K / (ms * 1e-3), "unit": "tiles/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": workload_config(B),
            "e2e": e2e_entry(world, B, K, e2e_ms, e2e_f32_ms),
            "gpu_launches": 17 * K,
            "roofline": {"kernel": "igemm_kernel<384, EPI_HEADS> (fused heads)", "bound": "tensor",
                         "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                         "traffic": (traffic * B // 64) if traffic is not None else None,
                         "traffic_source": ("dram read + write per launch from the committed ncu --set full capture %s, scaled to "
                                            "this batch; not re-measured in this run" % traffic_src) if traffic is not None else None,
                         "peak_source": peaks["source"] + (" sustained bf16 (timed region %.2f s >= 1 s)" % (ms * 1e-3) if long_region
                                                           else " burst bf16 (timed region %.2f s < 1 s: burst clocks)" % (ms * 1e-3)),
                         "frac_of_burst": ach / peaks["bf16_tflops"], "frac_of_sustained": ach / peaks["bf16_tflops_sustained"],
                         "ms_per_launch": heads_ms},
            "step_tflops": FLOPS_PER_TILE * B / (ms / K * 1e-3) / 1e12,
            "precision_plans": {
                "timed": "mixed: the bf16 model (weights rounded to bf16, carried in fp16 containers) x fp16 activations; "
                         "rel-RMS vs fp32 4.3e-3 / 5.4e-3 / 8.7e-3 (heat / regr / offset, profiles/accuracy_r02.json)",
                "bf16": {"value": world * B / (other_ms["bf16"] * 1e-3), "unit": "tiles/s", "ms_per_step": other_ms["bf16"],
                         "note": "weights and activations bf16: 5.9e-3 / 7.8e-3 / 1.26e-2"},
                "fp16": {"value": world * B / (other_ms["fp16"] * 1e-3), "unit": "tiles/s", "ms_per_step": other_ms["fp16"],
                         "note": "weights and activations fp16: 7e-4 / 1e-3 / 1.6e-3"}},
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
                             "last_loss": train_loss, **train_extra}
        if gpu_reference is not None:
            line["gpu_reference"] = gpu_reference
        if world == 1 and not args.no_cpu_baseline:
            rate, spi, iters, cores = cpu_oracle_rate(B, 12.0)
            line["cpu_baseline"] = {"value": rate, "unit": "tiles/s", "cores": cores, "kind": "port",
                                    "sample": "%d iterations x %d tiles (in chunks of 8), infer+decode, fp32 ATen on %d threads"
                                              % (iters, B, cores)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
