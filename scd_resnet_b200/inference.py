"""Batched tile inference + decode, the path BASELINE config 2 measures (test.py:95-112 per batch).

TileDetector keeps the packed weights, the activation workspace and the output buffers resident, and
runs   H2D copy -> scd_resnet_infer -> scd_decode_topk -> D2H copy   with uploads, kernels and downloads on
three streams so that the transfer of batch i+1 overlaps the kernels of batch i.
"""
import torch

from . import ops, weights


class TileDetector:
    """model: a CenterNetResidual in eval mode (or a state_dict).  batch: tiles per launch.
    precision: operand formats of the tensor-core path (weights.PRECISIONS): "mixed" (default: the bf16 model's weights
    x fp16 activations, within 1e-2 of the fp32 reference on every head), "bf16" (both bf16), "fp16" (both fp16)."""

    def __init__(self, model_or_sd, batch, device=None, K=100, height=512, width=512, precision=None):
        self.precision = precision or weights.DEFAULT_PRECISION
        self.fmt, self.wdtype = weights.precision_spec(self.precision)
        self.fp16 = self.fmt == 1
        sd = model_or_sd.state_dict() if hasattr(model_or_sd, "state_dict") else model_or_sd
        sd = {k.replace("module.", "", 1) if k.startswith("module.") else k: v for k, v in sd.items()}
        self.device = torch.device(device if device is not None else "cuda")
        self.batch, self.K, self.h, self.w = batch, K, height, width
        with torch.cuda.device(self.device):
            self.depth, self.dims, self.kdims = weights.arch_of(sd)      # numLayers, widths, kernel-level widths
            self.blob = weights.pack_infer_blob(sd, self.device, self.wdtype)
            self.workspace = torch.empty(ops.lib.scd_resnet_workspace_bytes(self.depth, ops._dims_arg(self.kdims), batch,
                                                                            height, width),
                                         dtype=torch.uint8, device=self.device)
            hw = (height // 4, width // 4)
            self.maps = (torch.empty(batch, 1, *hw, device=self.device), torch.empty(batch, 4, *hw, device=self.device),
                         torch.empty(batch, 2, *hw, device=self.device))
            self.copy_stream = torch.cuda.Stream(self.device)        # host -> device
            self.d2h_stream = torch.cuda.Stream(self.device)         # device -> host (own stream: a result copy that
            self.compute_stream = torch.cuda.Stream(self.device)     # waits for batch i must not block the upload of i+1)
            self.dev_in = [torch.empty(batch, 1, height, width, device=self.device) for _ in range(2)]
            self.in_ready = [torch.cuda.Event() for _ in range(2)]
            self.in_free = [torch.cuda.Event() for _ in range(2)]
            self.out_ready = [torch.cuda.Event() for _ in range(2)]
            self.dev_in_u8 = None                                    # grey-byte tiles: allocated on first use
            self._host_ring = None
        self.launches_per_batch = len(ops.resnet_conv_specs(self.depth, self.kdims)) + 3     # stem + igemms + heads + decode

    def detect_device(self, x, stage_events=None):
        """x (B,1,H,W) f32 on the device -> (10,B,K) f32 planes on the device (current stream)."""
        b = x.shape[0]
        heat, regr, off = [m[:b] for m in self.maps]
        ops.resnet_infer(x, self.blob, self.depth, self.kdims, self.workspace, (heat, regr, off), stage_events,
                         fmt=self.fmt)
        return ops.decode_topk(heat, regr, off, K=self.K, planes=True)[6]

    def detect_host(self, host_batches):
        """host_batches: list of pinned (B,1,H,W) host tensors, float32 (normalised tiles, what the network consumes) or
        uint8 (grey values as the slide holds them: a quarter of the bytes over the host link; the per-tile fp64
        normalisation of test.py:89 then runs on the device, scd_tiles_normalize_u8).  Returns a list of (10,B,K) host
        tensors.

        Copies run on copy_stream, kernels on compute_stream, two buffers each way."""
        n = len(host_batches)
        if any(hb.dtype == torch.uint8 for hb in host_batches) and self.dev_in_u8 is None:
            with torch.cuda.device(self.device):
                self.dev_in_u8 = [torch.empty(self.batch, 1, self.h, self.w, dtype=torch.uint8, device=self.device)
                                  for _ in range(2)]
        if self._host_ring is None or self._host_ring.shape[0] < n:
            self._host_ring = torch.empty(n, 10, self.batch, self.K, pin_memory=True)
        results = []
        with torch.cuda.device(self.device):
            # The buffers (and the weights) were allocated / written on the caller's stream: the side streams start behind
            # it.  (Without this the first call after construction could overlap the tail of the weight packing, and a
            # freshly allocated buffer could alias a block whose last use on the caller's stream was still running.)
            cur = torch.cuda.current_stream(self.device)
            for side in (self.copy_stream, self.compute_stream, self.d2h_stream):
                side.wait_stream(cur)
            self.t_first = torch.cuda.Event(enable_timing=True)
            self.t_last = torch.cuda.Event(enable_timing=True)
            self.t_first.record(self.copy_stream)             # device-side bracket of the whole call
            for i, hb in enumerate(host_batches):
                s = i & 1
                b = hb.shape[0]
                with torch.cuda.stream(self.copy_stream):
                    if i >= 2:
                        self.copy_stream.wait_event(self.in_free[s])      # batch i-2 has consumed this buffer
                    u8 = hb.dtype == torch.uint8
                    (self.dev_in_u8 if u8 else self.dev_in)[s][:b].copy_(hb, non_blocking=True)
                    self.in_ready[s].record(self.copy_stream)
                with torch.cuda.stream(self.compute_stream):
                    self.compute_stream.wait_event(self.in_ready[s])
                    if u8:
                        ops.tiles_normalize_u8(self.dev_in_u8[s][:b], out=self.dev_in[s])
                    planes = self.detect_device(self.dev_in[s][:b])
                    self.in_free[s].record(self.compute_stream)
                    self.out_ready[s].record(self.compute_stream)
                with torch.cuda.stream(self.d2h_stream):
                    self.d2h_stream.wait_event(self.out_ready[s])
                    out = self._host_ring[i, :, :b]
                    out.copy_(planes, non_blocking=True)
                    planes.record_stream(self.d2h_stream)
                    results.append(out)
            self.t_last.record(self.d2h_stream)
            self.d2h_stream.synchronize()
        return results
