"""Training step of centerOffsetRes10 (and the full-width Res18 / Res34 plugins) on hand-written sm_100a kernels.

Mirrors NetworkFactory.train (ref: models/networkFactory.py:257-263): zero_grad -> forward with
batch-statistics BatchNorm (ref: models/backbones/residuals.py:312-334) -> CenterNetLoss
(ref: models/centerNetOffset.py:182-217) -> backward -> Adam (torch defaults, :80-82).

TrainEngine owns four flat fp32 buffers: P (master parameters; the module's nn.Parameters are re-bound as
views into it, so state_dict / checkpoints / DDP-style broadcasts keep working), G (gradients, in the layouts
the wgrad kernels produce), M and V (Adam moments), plus WB, the bf16 GEMM-operand copies of the weights
(forward and data-gradient layouts) that one gather kernel refreshes after every optimiser step.
Orchestration is host Python over the C ABI; every device operation is one of this repo's kernels
(besides memsets and, for more than one rank, NCCL all-reduces).
"""
import torch

from . import ops, weights
from . import train_ops as T
from ._lib import ScdError

_HEADS = (("heatmap", 0, 1), ("regr", 1, 4), ("offset", 5, 2))
# BatchNorm batch statistics from the conv's store epilogue (scd_conv_igemm_fwd_bn); "0" = separate statistics pass
_CONV_BN_STATS = __import__("os").environ.get("SCD_CONV_BN_STATS", "1") != "0"
# the stem's pool backward + BatchNorm backward without the intermediate dy0 (scd_stem_bn_pool_bwd); "0" = separate kernels
_STEM_BWD_FUSED = __import__("os").environ.get("SCD_STEM_BWD_FUSED", "1") != "0"


def _block_list(depth, dims):
    """(block prefix, cin, cout, stride) of every BasicBlock (ref: ResNet.makeLayer, residuals.py:256-271)."""
    out, cin = [], dims[0]
    for li in range(1, 5):
        c = dims[li]
        for b in range(weights.BLOCKS[depth][li - 1]):
            out.append(("layer%d.%d" % (li, b), cin, c, 2 if (li > 1 and b == 0) else 1))
            cin = c
    return out


def _numel(shape):
    n = 1
    for d in shape:
        n *= d
    return n


def _deconv_list(dims):
    out, cin = [], dims[4]
    for i in range(3):
        out.append(("deconvolutionLayers.%d" % (3 * i), "deconvolutionLayers.%d" % (3 * i + 1), cin, dims[5 + i]))
        cin = dims[5 + i]
    return out


class TrainEngine:
    def __init__(self, module, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, regr_w=0.1, off_w=0.1, process_group=None,
                 peer_stats=True, stat_group="same", dense_heads=False):
        """process_group: ranks whose gradients this engine averages itself (None: it does not; a DistributedDataParallel
        wrapper around the module may).  stat_group: ranks that share BatchNorm batch statistics (= SyncBatchNorm);
        "same" = process_group.  dense_heads: also lay out the buffers of the dense heads backward, used when the
        upstream gradients of regr / offset arrive as dense maps (the autograd route with a foreign loss)."""
        self.module = module
        self.lr, self.betas, self.eps = lr, betas, eps
        self.regr_w, self.off_w = regr_w, off_w
        self.group = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.stat_group = process_group if isinstance(stat_group, str) else stat_group
        self.stat_world = torch.distributed.get_world_size(self.stat_group) if self.stat_group is not None else 1
        self.dense_heads = dense_heads
        self.step_count = 0
        named = dict(module.named_parameters())
        dev = next(module.parameters()).device
        if dev.type != "cuda":
            raise ScdError("TrainEngine needs the module on a CUDA device")
        self.dev = dev
        depth, dims, kd = weights.arch_of(module.state_dict())
        self.depth, self.dims, self.kd = depth, list(dims), list(kd)       # kd: the widths the kernels run at (>= 64)
        self.blocks, self.deconvs = _block_list(depth, kd), _deconv_list(kd)
        # ---- flat parameter buffer; head 1x1 weights / biases grouped so the kernels see (7,128), (7), (384).
        # A narrow network (half / quarter-width plugins) is trained zero-padded to the kernels' widths: per-channel
        # vectors (BatchNorm affine, head biases) and the heads' 1x1 rows are ALLOCATED at the padded width and the
        # module's parameter is a view of the leading part; conv weights stay in their own shape and are padded by the
        # operand gather (structural zeros) / read back through the gradient map.  A padded channel has zero weights,
        # gamma = beta = 0, hence activation 0 and gradient 0 everywhere, and Adam leaves exact zeros in place.
        order = [k for k in named if not (k.split(".")[0] in ("heatmap", "regr", "offset"))]
        order += [h + ".0.weight" for h, _, _ in _HEADS] + [h + ".0.bias" for h, _, _ in _HEADS]
        order += [h + ".2.weight" for h, _, _ in _HEADS] + [h + ".2.bias" for h, _, _ in _HEADS]
        assert sorted(order) == sorted(named)
        self.alloc = {}                                   # key -> allocated shape inside P
        # BatchNorm2d, or SyncBatchNorm after torch.nn.SyncBatchNorm.convert_sync_batchnorm (ref: networkFactory.py:133)
        bn_modules = {n: m for n, m in module.named_modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)}
        for k in order:
            shape = tuple(named[k].shape)
            prefix, leaf = k.rsplit(".", 1)
            if prefix in bn_modules:
                shape = (weights.pad_width(shape[0]),)
            elif k.split(".")[0] in ("heatmap", "regr", "offset"):
                if k.endswith(".0.bias"):
                    shape = (128,)
                elif k.endswith(".2.weight"):
                    shape = (shape[0], 128)
            self.alloc[k] = shape
        self.off, n = {}, 0
        for k in order:
            self.off[k] = n
            n += _numel(self.alloc[k])
        self.n_params = n
        self.P = torch.zeros(n, dtype=torch.float32, device=dev)
        for k in order:
            p = named[k]
            view = self._param_view(self.P, k, p.shape)
            view.copy_(p.data)
            p.data = view                                  # the module now lives in the flat buffer
        for prefix, m in bn_modules.items():               # running statistics at the padded width, too
            C, Cp = m.running_mean.shape[0], weights.pad_width(m.running_mean.shape[0])
            if Cp != C:
                rm, rv = torch.zeros(Cp, device=dev), torch.ones(Cp, device=dev)
                rm[:C].copy_(m.running_mean)
                rv[:C].copy_(m.running_var)
                m.running_mean.data, m.running_var.data = rm[:C], rv[:C]
        self._pending, self._reduced_from = [], 0          # gradient all-reduces in flight / lowest G index covered
        self.M = torch.zeros_like(self.P)
        self.V = torch.zeros_like(self.P)
        self._build_layouts(named)
        self.zero_bias = torch.zeros(512, dtype=torch.float32, device=dev)
        self.refresh_operands()
        # SyncBatchNorm statistics: 2 x C fp64 sums per BatchNorm, forward and backward.  Over NVLink peer memory when
        # the platform allows it (falls back to NCCL all-reduces otherwise; `peer_reason` says why).
        self.peer, self.peer_reason = None, ""
        self._versions = self._param_versions()
        if self.stat_world > 1 and peer_stats:
            from .dist import PeerAllReduce
            pa = PeerAllReduce(self.stat_group, dev, cap=1024)
            if pa.available:
                self.peer = pa
            else:
                self.peer_reason = pa.reason
        if hasattr(module, "_engine"):
            module._engine = self                          # the module's train-mode forward runs through this engine

    # ------------------------------------------------------------------ layouts
    def _build_layouts(self, named):
        """gmap: parameter element -> index in G;  wmap: element of WB -> index in P (or -1)."""
        gmap = torch.empty(self.n_params, dtype=torch.int64)
        self.g_off, g_n = {}, 0
        wparts, self.wb_off, w_n = [], {}, 0

        def g_alloc(key, floats):
            nonlocal g_n
            self.g_off[key] = g_n
            g_n += (floats + 63) // 64 * 64
            return self.g_off[key]

        def pidx(key, padded=None):                       # 1-based indices of a parameter inside P, zero-padded to `padded`
            p = named[key]
            assert self.alloc[key] == tuple(p.shape)       # conv weights: allocated in their own shape
            t = (torch.arange(p.numel(), dtype=torch.int64) + self.off[key] + 1).view(p.shape)
            return t if padded is None else weights._pad(t, padded)

        def w_alloc(key, index_tensor):
            nonlocal w_n
            flat = index_tensor.reshape(-1) - 1           # zeros (structural) -> -1
            self.wb_off[key] = (w_n, tuple(index_tensor.shape))
            w_n += (flat.numel() + 127) // 128 * 128
            wparts.append((self.wb_off[key][0], flat))

        def plain(key):
            n = _numel(self.alloc[key])
            base = g_alloc(key, n)
            gmap[self.off[key]:self.off[key] + n] = base + torch.arange(n)

        def conv(key, kind, cin, cout, fwd=True):
            """cin, cout: the (padded) widths the kernels run this conv at."""
            real = tuple(named[key].shape)
            padded = ((cin, cout) if kind == 3 else (cout, cin)) + real[2:]
            base = g_alloc(key, T.conv_wgrad_floats(kind, cin, cout))
            gi = weights.wgrad_index(padded, kind)[tuple(slice(0, d) for d in real)]
            gmap[self.off[key]:self.off[key] + named[key].numel()] = (base + gi).reshape(-1)
            if fwd:
                w_alloc(key + ":fwd", weights.layout_fwd(pidx(key, padded), kind))
            return padded

        def bn(prefix):
            plain(prefix + ".weight")
            plain(prefix + ".bias")

        # stem
        c0 = named["preprocess.0.weight"].shape[0]
        base = g_alloc("preprocess.0.weight", T.conv_wgrad_floats(4, 64, 64))
        gmap[self.off["preprocess.0.weight"]:self.off["preprocess.0.weight"] + c0 * 49] = \
            (base + weights.wgrad_index((64, 1, 7, 7), 4)[:c0]).reshape(-1)
        w_alloc("preprocess.0.weight:fwd", weights.layout_stem(pidx("preprocess.0.weight", (64, 1, 7, 7))))
        bn("preprocess.1")
        for p, cin, cout, stride in self.blocks:
            pad1 = conv(p + ".conv1.weight", 0 if stride == 1 else 1, cin, cout)
            bn(p + ".bn1")
            pad2 = conv(p + ".conv2.weight", 0, cout, cout)
            bn(p + ".bn2")
            w_alloc(p + ".conv2.weight:dgrad", weights.layout_dgrad(pidx(p + ".conv2.weight", pad2), 0))
            if stride == 1:
                w_alloc(p + ".conv1.weight:dgrad", weights.layout_dgrad(pidx(p + ".conv1.weight", pad1), 0))
            else:
                padd = conv(p + ".downsample.0.weight", 2, cin, cout)
                bn(p + ".downsample.1")
                w_alloc(p + ".conv1.weight:dgrad",
                        weights.layout_dgrad(pidx(p + ".conv1.weight", pad1), 1, pidx(p + ".downsample.0.weight", padd)))
        for ck, bk, cin, cout in self.deconvs:
            padc = conv(ck + ".weight", 3, cin, cout)
            bn(bk)
            w_alloc(ck + ".weight:dgrad", weights.layout_dgrad(pidx(ck + ".weight", padc), 3))
        # heads: the three 3x3 convs run as one conv with 384 output channels in the forward pass.  Backward: the
        # heat head (dense gradient) goes through the tensor-core wgrad / dgrad with 128 channels; the regr and
        # offset heads have a hidden gradient at the object pixels only (csrc/heads_sparse.cu), their weight
        # gradient is laid out [tap][co][ci] with co = 0..127 regr, 128..255 offset.
        hc = self.kd[7]                                                                 # channels of the heads' input
        hd = named["heatmap.0.weight"].shape[0]                                         # head width: 128, or 64 (h / q)
        ci_real = named["heatmap.0.weight"].shape[1]
        w3_idx = torch.cat([pidx(h + ".0.weight", (128, hc, 3, 3)) for h, _, _ in _HEADS], 0)   # (384,hc,3,3) of P indices
        # heat head weight gradient: wgrad kind 5 (dz shifted, N = the heads' input width) when that is wider than 128
        self.heat_wgrad_kind = 5 if hc > 128 else 0
        base = g_alloc("heads.w3h", T.conv_wgrad_floats(self.heat_wgrad_kind, hc, 128))
        k = "heatmap.0.weight"
        gmap[self.off[k]:self.off[k] + named[k].numel()] = \
            (base + weights.wgrad_index((128, hc, 3, 3), self.heat_wgrad_kind)[:hd, :ci_real]).reshape(-1)
        base = g_alloc("heads.w3s", 9 * 256 * hc)
        co, ci, r, s_ = torch.meshgrid(torch.arange(hd), torch.arange(ci_real), torch.arange(3), torch.arange(3),
                                       indexing="ij")
        for i, k in enumerate(("regr.0.weight", "offset.0.weight")):
            gmap[self.off[k]:self.off[k] + named[k].numel()] = \
                (base + ((r * 3 + s_) * 256 + (co + 128 * i)) * hc + ci).reshape(-1)
        w_alloc("heads.w3:fwd", weights.layout_fwd(w3_idx, 0))
        w_alloc("heads.w3h:dgrad", weights.layout_dgrad(w3_idx[:128], 0))
        b3 = g_alloc("heads.b3", 384)
        w1 = g_alloc("heads.w1", 7 * 128)
        b1 = g_alloc("heads.b1", 7)
        for i, (h, j0, nj) in enumerate(_HEADS):
            gmap[self.off[h + ".0.bias"]:self.off[h + ".0.bias"] + 128] = b3 + i * 128 + torch.arange(128)
            gmap[self.off[h + ".2.weight"]:self.off[h + ".2.weight"] + nj * 128] = w1 + j0 * 128 + torch.arange(nj * 128)
            gmap[self.off[h + ".2.bias"]:self.off[h + ".2.bias"] + nj] = b1 + j0 + torch.arange(nj)
        self.n_grads = g_n                                                              # what a gradient all-reduce covers
        self._reduced_from = g_n
        self.gmap = gmap.to(torch.int32).to(self.dev)
        self.gmap_dense = None
        if self.dense_heads:
            # dense heads backward: one wgrad (kind 0, 384 output channels) / one dgrad over all three heads; the slot
            # sits behind the all-reduced range and a second gradient map reads the 3x3 weights from it
            base = g_alloc("heads.w3d", T.conv_wgrad_floats(0, hc, 384))
            gd = gmap.clone()
            idx384 = weights.wgrad_index((384, hc, 3, 3), 0)
            for i, (h, _, _) in enumerate(_HEADS):
                k = h + ".0.weight"
                gd[self.off[k]:self.off[k] + named[k].numel()] = (base + idx384[i * 128:i * 128 + hd, :ci_real]).reshape(-1)
            self.gmap_dense = gd.to(torch.int32).to(self.dev)
            w_alloc("heads.w3:dgrad", weights.layout_dgrad(w3_idx, 0))
        self.G = torch.zeros(g_n, dtype=torch.float32, device=self.dev)
        wmap = torch.full((w_n,), -1, dtype=torch.int64)
        for o, flat in wparts:
            wmap[o:o + flat.numel()] = flat
        self.wmap = wmap.to(torch.int32).to(self.dev)
        self.WB = torch.empty(w_n, dtype=torch.bfloat16, device=self.dev)

    def _param_view(self, flat, key, shape):
        """The view of parameter `key` (reference shape `shape`) inside a flat buffer laid out like P."""
        a = self.alloc[key]
        t = flat[self.off[key]:self.off[key] + _numel(a)].view(a)
        if a == tuple(shape):
            return t
        if len(a) == 1:                                   # per-channel vector allocated at the padded width
            return t[:shape[0]]
        return t[:, :shape[1]].unsqueeze(-1).unsqueeze(-1)       # head 1x1 weight: rows of 128, (nj, hd, 1, 1) in the module

    def wb(self, key):
        o, shape = self.wb_off[key]
        n = 1
        for s in shape:
            n *= s
        return self.WB[o:o + n]

    def g(self, key, n=None):
        o = self.g_off[key]
        return self.G[o:o + n] if n is not None else self.G[o:]

    def p(self, key):
        return dict(self.module.named_parameters())[key].data

    def refresh_operands(self):
        T.gather_cast_bf16(self.P, self.wmap, self.WB)
        self._versions = self._param_versions()

    def _param_versions(self):
        return tuple(p._version for p in self.module.parameters())

    def refresh_if_changed(self):
        """The 16-bit operand copies follow the fp32 masters: any in-place update of a parameter since the last refresh
        (a torch optimizer stepping on the views, load_state_dict, a broadcast) triggers one gather kernel."""
        if self._versions != self._param_versions():
            self.refresh_operands()

    def owns(self, module=None):
        """True while every parameter of the module still is the view into P this engine bound it to (module.to(),
        .half() or a re-assigned .data break that; the caller then builds a new engine)."""
        base = self.P.data_ptr()
        for k, p in (module or self.module).named_parameters():
            if k not in self.off or p.data_ptr() != base + 4 * self.off[k] or p.dtype != torch.float32:
                return False
        return True

    # ------------------------------------------------------------------ SyncBatchNorm hooks
    def _allreduce_stats(self, sums, pixels):
        """NCCL route of the statistics exchange (used when peer memory is unavailable; with it, the exchange runs inside
        the BatchNorm reduction kernels, csrc/bn.cu)."""
        if self.stat_world == 1:
            return float(pixels) if pixels is not None else None
        if self.peer is not None:
            self.peer(sums)
        else:
            torch.distributed.all_reduce(sums, group=self.stat_group)
        return float(pixels) * self.stat_world if pixels is not None else None

    def _sync_kw(self):
        fused = T._bn_fused(self.stat_world) and self.peer is not None
        return {"all_reduce": self._allreduce_stats if (self.stat_world > 1 and not fused) else None,
                "peer": self.peer if fused else None, "world": self.stat_world}

    # ------------------------------------------------------------------ forward + backward
    def _bn(self, z, prefix, residual=None, relu=True):
        m = self.module.get_submodule(prefix)
        return T.bn_forward(z, m.weight.data, m.bias.data, m.running_mean, m.running_var, m.num_batches_tracked,
                            residual, relu, **self._sync_kw())

    def _bn_bwd(self, da, a, z, ctx, prefix, want_dy=False, relu_from_z=False):
        return T.bn_backward(da, a, z, ctx, want_dy, self.g(prefix + ".weight"), self.g(prefix + ".bias"),
                             relu_from_z=relu_from_z, **self._sync_kw())

    def _conv(self, kind, x, key, cout):
        return ops.conv_igemm_fwd(kind, x, self.wb(key + ":fwd"), self.zero_bias[:cout], None, False)

    def _conv_bn(self, kind, x, key, cout, prefix, residual=None, relu=True):
        """conv -> train-mode BN (+residual)(+ReLU) -> (z, a, ctx).  The 3x3 / 4x4 stages accumulate the batch statistics
        in the conv's own store epilogue; the 1x1 downsample (one k-block per tile, epilogue-bound already) keeps the
        separate statistics pass.  SCD_CONV_BN_STATS=0 restores the separate pass everywhere."""
        if kind == 2 or not _CONV_BN_STATS:
            z = self._conv(kind, x, key, cout)
            a, ctx = self._bn(z, prefix, residual=residual, relu=relu)
            return z, a, ctx
        m = self.module.get_submodule(prefix)
        return T.conv_bn_forward(kind, x, self.wb(key + ":fwd"), self.zero_bias[:cout], cout, m.weight.data, m.bias.data,
                                 m.running_mean, m.running_var, m.num_batches_tracked, residual, relu, **self._sync_kw())

    def forward(self, x, keep=True):
        """Train-mode forward (batch-statistics BatchNorm, running statistics updated) on x (B,1,H,W) f32 CUDA.
        Returns ((heat logits, regr, offset) NCHW f32, tape); tape = what backward() needs, or None when not `keep`
        (validation in train mode under no_grad, ref: models/networkFactory.py:265-271 with the model left in train())."""
        mod = self.module
        self.refresh_if_changed()
        z0, col0 = T.stem_conv_train(x, self.wb("preprocess.0.weight:fwd"))
        _, ctx0 = self._stem_bn(z0, mod.preprocess[1])
        a, argmax0 = T.stem_bn_relu_pool(z0, ctx0["stat"], want_argmax=keep)
        tape = []
        for p, cin, cout, stride in self.blocks:
            a_in = a
            z1, a1, c1 = self._conv_bn(0 if stride == 1 else 1, a_in, p + ".conv1.weight", cout, p + ".bn1")
            if stride == 1:
                skip, zd, cd = a_in, None, None
            else:
                zd, skip, cd = self._conv_bn(2, a_in, p + ".downsample.0.weight", cout, p + ".downsample.1", relu=False)
            z2, a, c2 = self._conv_bn(0, a1, p + ".conv2.weight", cout, p + ".bn2", residual=skip)
            if keep:
                tape.append((p, cin, cout, stride, a_in, z1, a1, c1, z2, c2, zd, cd, a))
        dtape = []
        for ck, bk, cin, cout in self.deconvs:
            a_in = a
            z, a, c = self._conv_bn(3, a_in, ck + ".weight", cout, bk)
            if keep:
                dtape.append((ck, bk, cin, cout, a_in, z, c, a))
        e3 = a
        b3 = self.P[self.off["heatmap.0.bias"]:self.off["heatmap.0.bias"] + 384]
        w1 = self.P[self.off["heatmap.2.weight"]:self.off["heatmap.2.weight"] + 7 * 128]
        b1 = self.P[self.off["heatmap.2.bias"]:self.off["heatmap.2.bias"] + 7]
        heat, regr, off, hidden = T.heads_fwd_train(e3, self.wb("heads.w3:fwd"), b3, w1, b1)
        if not keep:
            return (heat, regr, off), None
        return (heat, regr, off), {"z0": z0, "col0": col0, "ctx0": ctx0, "argmax0": argmax0, "blocks": tape,
                                   "deconvs": dtape, "e3": e3, "hidden": hidden, "w1": w1}

    def backward(self, tape, d_heat, sparse=None, dense=None):
        """Backward pass from the heads' output gradients; weight gradients land in self.G (wgrad layouts).
        d_heat (B,1,H,W) f32 and EITHER sparse = (d_obj (B,30,6), mask, idx): the masked-L1 gradients of regr / offset in
        per-object form (ops.centernet_loss_sparse), OR dense = (d_regr (B,4,H,W), d_off (B,2,H,W)) (needs
        dense_heads=True)."""
        for w in self._pending:                            # a backward pass that was not followed by optimizer_step
            w.wait()
        self._pending, self._reduced_from = [], self.n_grads
        self.G.zero_()
        e3, hidden, w1 = tape["e3"], tape["hidden"], tape["w1"]
        hc = self.kd[7]
        if sparse is not None:
            d_obj, mask, gidx = sparse
            d_hh, dh_obj = T.heads_bwd_sparse(d_heat, d_obj, mask, gidx, hidden, w1, self.g("heads.w1"),
                                              self.g("heads.b1"), self.g("heads.b3"))
            T.conv_wgrad(self.heat_wgrad_kind, e3, d_hh, hc, 128, self.g("heads.w3h"))
            T.heads_wgrad_sparse(e3, dh_obj, mask, gidx, self.g("heads.w3s"))
            da = T.conv_dgrad(0, d_hh, self.wb("heads.w3h:dgrad"), self.zero_bias[:hc], hc)
            T.heads_dgrad_sparse(dh_obj, mask, gidx, self.wb("heads.w3:fwd"), da)
        else:
            if not self.dense_heads:
                raise ScdError("TrainEngine: dense heads backward needs dense_heads=True")
            d_regr, d_off = dense
            d_hidden = T.heads_bwd(d_heat, d_regr, d_off, hidden, w1, self.g("heads.w1"), self.g("heads.b1"),
                                   self.g("heads.b3"))
            T.conv_wgrad(0, e3, d_hidden, hc, 384, self.g("heads.w3d"))
            da = T.conv_dgrad(0, d_hidden, self.wb("heads.w3:dgrad"), self.zero_bias[:hc], hc)
        for ck, bk, cin, cout, a_in, z, c, a_out in reversed(tape["deconvs"]):
            dz, _ = self._bn_bwd(da, None, z, c, bk, relu_from_z=True)       # conv -> BN -> ReLU, no residual
            T.conv_wgrad(3, a_in, dz, cin, cout, self.g(ck + ".weight"))
            da = T.conv_dgrad(3, dz, self.wb(ck + ".weight:dgrad"), self.zero_bias[:cin], cin)
        self._reduce_async(self.g_off[self.deconvs[0][0] + ".weight"], self.n_grads)          # deconvs + heads are final
        for p, cin, cout, stride, a_in, z1, a1, c1, z2, c2, zd, cd, a_out in reversed(tape["blocks"]):
            dz2, dy = self._bn_bwd(da, a_out, z2, c2, p + ".bn2", want_dy=True)
            T.conv_wgrad(0, a1, dz2, cout, cout, self.g(p + ".conv2.weight"))
            da1 = T.conv_dgrad(0, dz2, self.wb(p + ".conv2.weight:dgrad"), self.zero_bias[:cout], cout)
            dz1, _ = self._bn_bwd(da1, None, z1, c1, p + ".bn1", relu_from_z=True)
            if stride == 1:
                T.conv_wgrad(0, a_in, dz1, cin, cout, self.g(p + ".conv1.weight"))
                da = T.conv_dgrad(0, dz1, self.wb(p + ".conv1.weight:dgrad"), self.zero_bias[:cin], cin, add=dy)
            else:
                dzd, _ = self._bn_bwd(dy, None, zd, cd, p + ".downsample.1")
                T.conv_wgrad(1, a_in, dz1, cin, cout, self.g(p + ".conv1.weight"))
                T.conv_wgrad(2, a_in, dzd, cin, cout, self.g(p + ".downsample.0.weight"))
                da = T.conv_dgrad(1, dz1, self.wb(p + ".conv1.weight:dgrad"), self.zero_bias[:cin], cin, dz2=dzd)
            # Everything from this block's first gradient upwards is final: start its all-reduce now, on NCCL's stream,
            # while the backward pass goes on below (layer3 + layer4 together, then layer2, then layer1; only the stem's
            # 3 K gradients are left for optimizer_step).
            if p in ("layer3.0", "layer2.0", "layer1.0"):
                self._reduce_async(self.g_off[p + ".conv1.weight"], self._reduced_from)
        if _STEM_BWD_FUSED:
            dz0 = T.stem_bn_pool_backward(tape["argmax0"], da, tape["z0"], tape["ctx0"], self.g("preprocess.1.weight"),
                                          self.g("preprocess.1.bias"), **self._sync_kw())
        else:
            dy0 = T.stem_pool_bwd(tape["argmax0"], da)
            dz0, _ = T.bn_backward(dy0, None, tape["z0"], tape["ctx0"], False, self.g("preprocess.1.weight"),
                                   self.g("preprocess.1.bias"), **self._sync_kw())
        T.conv_wgrad(4, tape["col0"], dz0, 64, 64, self.g("preprocess.0.weight"))

    def forward_backward(self, x, targets, sigmoid_inplace=False):
        """x (B,1,H,W) f32 CUDA; targets = [heat (B,1,128,128), mask (B,30), regr6 (B,30,6), idx (B,30)]
        (+ optionally the two device counters of ops.render_targets(with_npos=True)).
        Returns (losses f32[4] on the device, outputs dict); gradients land in self.G."""
        (heat, regr, off), tape = self.forward(x, keep=True)
        # ---- loss (forward + its own backward in one pass; L1 gradients in sparse, per-object form) -----------
        gt_heat, mask, regr6, gidx = targets[0], targets[1], targets[2], targets[3]
        counts = targets[4] if len(targets) > 4 else None          # [N_pos, mask.sum()] from render_targets(with_npos)
        losses, d_heat, d_obj = ops.centernet_loss_sparse(heat, regr, off, gt_heat, mask, regr6, gidx,
                                                          self.regr_w, self.off_w, npos=counts,
                                                          sigmoid_inplace=sigmoid_inplace)
        self.backward(tape, d_heat, sparse=(d_obj, mask, gidx))
        return losses, {"heatmap": heat, "regr": regr, "offset": off}

    def _stem_bn(self, z0, bn0):
        """Batch statistics of the stem conv output; the normalisation itself is fused with ReLU + max-pool."""
        C = 64
        pixels = z0.numel() // C
        sums = torch.empty(2 * C + 1, dtype=torch.float64, device=self.dev)
        stat = torch.empty(4, C, dtype=torch.float32, device=self.dev)
        outs = (ops._ptr(stat[0]), ops._ptr(stat[1]), ops._ptr(stat[2]), ops._ptr(stat[3]))
        bnp = (ops._ptr(bn0.weight.data), ops._ptr(bn0.bias.data), ops._ptr(bn0.running_mean), ops._ptr(bn0.running_var),
               ops._ptr(bn0.num_batches_tracked))
        if T._bn_fused(self.stat_world) and (self.stat_world == 1 or self.peer is not None):
            count = float(pixels) * self.stat_world
            pa = self.peer.next_args() if (self.peer is not None and self.stat_world > 1) else T._NO_PEER
            ops.check(ops.lib.scd_bn_stats_finalize(ops._ptr(z0), pixels, C, ops._ptr(sums), *bnp, count, T.BN_MOMENTUM,
                                                    T.BN_EPS, *outs, *pa, ops._stream()), "scd_bn_stats_finalize")
        else:
            ops.check(ops.lib.scd_bn_stats(ops._ptr(z0), pixels, C, ops._ptr(sums), ops._stream()), "scd_bn_stats")
            count = self._allreduce_stats(sums[:2 * C], pixels)
            ops.check(ops.lib.scd_bn_finalize(ops._ptr(sums), *bnp, C, count, T.BN_MOMENTUM, T.BN_EPS, *outs, ops._stream()),
                      "scd_bn_finalize")
        return None, {"stat": stat, "count": count, "sums": sums}

    # ------------------------------------------------------------------ optimiser
    def _reduce_async(self, lo, hi):
        """Start the gradient all-reduce of G[lo:hi] (a range whose weight gradients are final) while the backward pass
        goes on: NCCL runs it on its own stream, ordered after what this stream has written so far."""
        if self.world > 1 and hi > lo:
            self._pending.append(torch.distributed.all_reduce(self.G[lo:hi], group=self.group, async_op=True))
            self._reduced_from = min(self._reduced_from, lo)

    def finish_reduce(self):
        """Complete the gradient all-reduce (sum over ranks) of G[0:n_grads).  The deconv + heads range and the layer3-4
        range were started during the backward pass; what is left here is the stem .. layer2 range (12 % of the
        parameters)."""
        if self.world > 1:
            self._reduce_async(0, min(self._reduced_from, self.n_grads))
            for w in self._pending:
                w.wait()                                   # this stream waits for the collectives, not the host
            self._pending, self._reduced_from = [], self.n_grads

    def optimizer_step(self):
        """DDP semantics (ref: models/networkFactory.py:134): gradients are averaged over ranks, then Adam."""
        self.finish_reduce()
        if self.peer is not None:
            self.peer.check()
        self.step_count += 1
        T.adam_step(self.P, self.M, self.V, self.G, self.gmap, self.step_count, self.lr, self.betas, self.eps,
                    1.0 / self.world)
        self.refresh_operands()

    def train_step(self, x, targets):
        """NetworkFactory.train: returns the device tensor (total, focal, size, offset)."""
        losses, _ = self.forward_backward(x, targets)
        self.optimizer_step()
        return losses

    def set_learning_rate(self, lr):
        self.lr = lr

    def grads_reference_layout(self, dense=False, scale=None):
        """{name: gradient in the parameter's own layout}: one gather kernel from the wgrad layouts of G into a fresh flat
        buffer laid out like P (what loss.backward() leaves in param.grad, ref: models/networkFactory.py:261).
        dense: the last backward ran the dense heads path; scale: optional device scalar multiplied in."""
        g = torch.empty_like(self.P)
        T.gather_f32(self.G, self.gmap_dense if dense else self.gmap, g, scale)
        return {k: self._param_view(g, k, p.shape) for k, p in self.module.named_parameters()}


