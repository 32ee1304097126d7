"""Whole-slide inference, the flow of analyseImages (ref: test.py:21-33, 41-142).

RGB or grey slide -> (grayscale) -> reflect pad + stride-384 512x512 tiles + per-tile fp64 normalise (on the device,
csrc/slide.cu) -> batched inference + decode -> keep score > 0.3 and map to slide coordinates (on the device,
scd_slide_merge) -> ordered concatenation.  Tiles are enumerated x-major then y like the reference; with several ranks
each takes a contiguous range of that list, i.e. a few tile COLUMNS: it uploads only the slide columns those read, in
per-column chunks on a copy stream that overlap the kernels of the previous column, and only the kept detection rows
are gathered.  The merge is the reference's: no cross-tile de-duplication (its margin filter is commented out,
test.py:132-134)."""
import numpy as np
import torch

from . import ops, dist as sdist

INPUTSIZE, PADDINGSIZE = 512, 64            # ref: test.py:16-17


def grayscale(rgb):
    """ref: test.py:21-33: OpenCV-style weights on the first three channels of an (H,W,C) array, rounded.
    A CUDA uint8 tensor is converted on the device (scd_grayscale_u8, bit exact); anything else on the host."""
    if isinstance(rgb, torch.Tensor) and rgb.is_cuda:
        return ops.grayscale(rgb)
    rgb = np.asarray(rgb)
    return np.round(0.1140 * rgb[:, :, 0] + 0.5870 * rgb[:, :, 1] + 0.2989 * rgb[:, :, 2])


def merge_detections(planes, height, width, threshold=0.3):
    """ref: test.py:103-140 on the host.  planes: (10, T, K) float32 (host) for ALL tiles in order.
    Returns an (n, 3) float64 array of [x, y, ratio] rows in the reference's order."""
    clip_h, clip_v, _, _, pad_tb, pad_lr = ops.slide_geometry(height, width)
    step = INPUTSIZE - 2 * PADDINGSIZE
    p32 = planes.float().numpy() if isinstance(planes, torch.Tensor) else np.asarray(planes, np.float32)
    p = p32.astype(np.float64)
    sc, cy, cx, minl, rad, offx, offy = p[0], p[2], p[3], p[6], p[7], p[8], p[9]
    # the reference compares the float32 scores with the Python scalar in float32 (`ctScores[item] > 0.3`, test.py:104)
    t_idx, k_idx = np.nonzero(p32[0] > np.float32(threshold))   # row-major: tile order, then rank inside the tile
    tx, ty = t_idx // clip_v, t_idx % clip_v                   # x-major then y (test.py:114-116)
    dminl = minl[t_idx, k_idx] * 4
    halo = rad[t_idx, k_idx] * 4
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = (halo - dminl) / (2 * dminl)
    gx = np.trunc(tx * step - pad_lr + cx[t_idx, k_idx] * 4 + offx[t_idx, k_idx])      # int() truncates
    gy = np.trunc(ty * step - pad_tb + cy[t_idx, k_idx] * 4 + offy[t_idx, k_idx])
    return np.stack([gx, gy, ratio], axis=1)


def _as_host_2d(gray):
    """(H,W[,C]) uint8 / float32 host tensor with unit innermost stride; uint8 stays uint8 (a quarter of the upload)."""
    g = gray if isinstance(gray, torch.Tensor) else torch.as_tensor(np.asarray(gray))
    if g.dtype != torch.uint8:
        g = g.to(torch.float32)
    return g.contiguous()


def analyse_slide(detector, gray, threshold=0.3, group=None, return_planes=True):
    """detector: inference.TileDetector.  gray: (H,W) grey values or (H,W,C>=3) uint8 RGB, a host array / tensor (pinned
    memory makes the strip uploads asynchronous) or a CUDA tensor.
    Returns (detections (n,3) float64 [x, y, ratio] in the reference's order, identical on every rank;
             planes (10,T,K) float32 on the host for all tiles, or None when not `return_planes`)."""
    dev = detector.device
    on_device = isinstance(gray, torch.Tensor) and gray.is_cuda
    g = gray if on_device else _as_host_2d(gray)
    rgb = g.dim() == 3
    if rgb and g.dtype != torch.uint8:
        raise ops.ScdError("an RGB slide must be uint8 (H,W,C)")
    h, w = int(g.shape[0]), int(g.shape[1])
    clip_h, clip_v = ops.slide_geometry(h, w)[:2]
    total = clip_h * clip_v
    dist_on = torch.distributed.is_available() and torch.distributed.is_initialized()
    rank = torch.distributed.get_rank(group) if dist_on else 0
    world = torch.distributed.get_world_size(group) if dist_on else 1
    begin, end = sdist.shard_range(total, rank, world)
    K, B = detector.K, detector.batch
    cap = max(1, -(-total // world)) * K                       # rows a rank can produce (every detection kept)
    planes_local = []
    with torch.cuda.device(dev):
        rows = torch.empty(cap, 3, dtype=torch.float64, device=dev)
        count = torch.zeros(1, dtype=torch.int32, device=dev)
        if end > begin:
            tx0, tx1 = begin // clip_v, (end - 1) // clip_v
            spans = [ops.slide_column_span(h, w, tx) for tx in range(tx0, tx1 + 1)]
            c_lo, c_hi = min(s[0] for s in spans), max(s[1] for s in spans)
            compute = torch.cuda.current_stream()
            if on_device:
                gg = ops.grayscale(g) if rgb else (g if g.dtype == torch.uint8 else g.float())
                strip, col0, ready = gg, 0, None
            else:
                # the column strip [c_lo, c_hi), uploaded per tile column on the detector's copy stream
                strip = torch.empty(h, c_hi - c_lo, dtype=torch.uint8 if g.dtype == torch.uint8 else torch.float32, device=dev)
                stage = torch.empty(h, c_hi - c_lo, g.shape[2], dtype=torch.uint8, device=dev) if rgb else None
                col0, ready, done_hi = c_lo, [], c_lo
                copy = detector.copy_stream
                copy.wait_stream(compute)                      # the buffers exist before the copies start
                strip.record_stream(copy)
                if stage is not None:
                    stage.record_stream(copy)
                for lo, hi in spans:
                    a, b = max(done_hi, lo), max(done_hi, hi)
                    if b > a:
                        if rgb:
                            c = g.shape[2]
                            ops.copy2d_h2d(stage.view(h, -1)[:, (a - c_lo) * c:(b - c_lo) * c],
                                           g.view(h, -1)[:, a * c:b * c], copy)
                        else:
                            ops.copy2d_h2d(strip[:, a - c_lo:b - c_lo], g[:, a:b], copy)
                        done_hi = b
                    ev = torch.cuda.Event()
                    ev.record(copy)
                    ready.append((ev, done_hi))
                gray_done = c_lo
            for b0 in range(begin, end, B):
                b1 = min(b0 + B, end)
                if ready is not None:
                    ev, upto = ready[(b1 - 1) // clip_v - tx0]
                    compute.wait_event(ev)
                    if rgb and upto > gray_done:               # grayscale of the columns that have just landed
                        ops.grayscale(stage[:, gray_done - c_lo:upto - c_lo], out=strip[:, gray_done - c_lo:upto - c_lo])
                        gray_done = upto
                tiles = ops.slide_tiles_strip(strip, h, w, col0, b0, b1)
                planes = detector.detect_device(tiles)
                ops.slide_merge(planes, b0, h, w, rows, count, threshold)
                if return_planes:
                    planes_local.append(planes)
        # ---- gather: only the kept rows travel (fixed-shape buffers: cap rows per rank + the counts)
        if world > 1:
            all_rows = torch.empty(world, cap, 3, dtype=torch.float64, device=dev)
            all_counts = torch.empty(world, dtype=torch.int32, device=dev)
            torch.distributed.all_gather_into_tensor(all_rows, rows, group=group)
            torch.distributed.all_gather_into_tensor(all_counts, count, group=group)
            counts = all_counts.cpu().tolist()
            host = all_rows.cpu().numpy()
            dets = np.concatenate([host[r, :min(counts[r], cap)] for r in range(world)], axis=0)
        else:
            n = min(int(count.item()), cap)
            dets = rows[:n].cpu().numpy()
        planes_all = None
        if return_planes:
            local = torch.cat(planes_local, dim=1) if planes_local else torch.empty(10, 0, K, device=dev)
            planes_all = sdist.gather_planes(local, total, group).cpu()
    return dets, planes_all
