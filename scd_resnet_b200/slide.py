"""Whole-slide inference, the flow of analyseImages (ref: test.py:41-142).

grey slide -> reflect pad + stride-384 512x512 tiles + per-tile fp64 normalise (on the device, csrc/slide.cu)
-> batched inference + decode -> keep score > 0.3 -> map to slide coordinates.  Tiles are enumerated x-major
then y like the reference; with several ranks each takes a contiguous range of tiles and the per-tile
detections are all-gathered before the (ordered, dedup-free) merge, which is identical on every rank."""
import numpy as np
import torch

from . import ops, dist as sdist

INPUTSIZE, PADDINGSIZE = 512, 64            # ref: test.py:16-17


def grayscale(rgb):
    """ref: test.py:21-33: OpenCV-style weights on the first three channels of an (H,W,3) array, rounded."""
    rgb = np.asarray(rgb)
    return np.round(0.1140 * rgb[:, :, 0] + 0.5870 * rgb[:, :, 1] + 0.2989 * rgb[:, :, 2])


def merge_detections(planes, height, width, threshold=0.3):
    """ref: test.py:103-140.  planes: (10, T, K) float32 (host) for ALL tiles in order.
    Returns an (n, 3) float64 array of [x, y, ratio] rows in the reference's order (no cross-tile dedup; the
    reference's margin filter is commented out, test.py:132-134)."""
    clip_h, clip_v, _, _, pad_tb, pad_lr = ops.slide_geometry(height, width)
    step = INPUTSIZE - 2 * PADDINGSIZE
    p = planes.double().numpy() if isinstance(planes, torch.Tensor) else np.asarray(planes, np.float64)
    sc, cy, cx, minl, rad, offx, offy = p[0], p[2], p[3], p[6], p[7], p[8], p[9]
    t_idx, k_idx = np.nonzero(sc > threshold)                  # row-major: tile order, then rank inside the tile
    tx, ty = t_idx // clip_v, t_idx % clip_v                   # x-major then y (test.py:114-116)
    dminl = minl[t_idx, k_idx] * 4
    halo = rad[t_idx, k_idx] * 4
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = (halo - dminl) / (2 * dminl)
    gx = np.trunc(tx * step - pad_lr + cx[t_idx, k_idx] * 4 + offx[t_idx, k_idx])      # int() truncates
    gy = np.trunc(ty * step - pad_tb + cy[t_idx, k_idx] * 4 + offy[t_idx, k_idx])
    return np.stack([gx, gy, ratio], axis=1)


def analyse_slide(detector, gray, threshold=0.3, group=None):
    """detector: inference.TileDetector.  gray: (H,W) array or tensor of grey values; uint8 input stays uint8 on
    its way to the device (a quarter of the upload), anything else goes as float32.
    Returns (detections (n,3) float64 [x, y, ratio], planes (10,T,K) float32 on the host)."""
    g = gray if isinstance(gray, torch.Tensor) else torch.as_tensor(np.asarray(gray))
    g = g.to(device=detector.device, dtype=torch.uint8 if g.dtype == torch.uint8 else torch.float32, non_blocking=True)
    h, w = g.shape
    clip_h, clip_v = ops.slide_geometry(h, w)[:2]
    total = clip_h * clip_v
    rank = torch.distributed.get_rank(group) if torch.distributed.is_initialized() else 0
    world = torch.distributed.get_world_size(group) if torch.distributed.is_initialized() else 1
    begin, end = sdist.shard_range(total, rank, world)
    out = []
    with torch.cuda.device(detector.device):
        for b0 in range(begin, end, detector.batch):
            b1 = min(b0 + detector.batch, end)
            tiles = ops.slide_tiles(g, b0, b1)
            out.append(detector.detect_device(tiles))
        local = torch.cat(out, dim=1) if out else torch.empty(10, 0, detector.K, device=detector.device)
        planes = sdist.gather_planes(local, total, group).cpu()
    return merge_detections(planes, h, w, threshold), planes
