"""Training loop with the surface of the reference's NetworkFactory (ref: models/networkFactory.py:36-302):
plugin lookup by module path, train / validate, learning-rate decay, snapshots.

What differs underneath: `train` is one TrainEngine step (forward with batch-statistics BatchNorm, fused
CenterNetLoss, backward, fused Adam, all on this repo's sm_100a kernels); with several ranks the engine
all-reduces the flat gradient buffer and the BatchNorm statistics over NCCL, which is the DDP + SyncBatchNorm
semantics of networkFactory.py:126-134.  Parameters are broadcast from rank 0 at start, as DDP does."""
import importlib
import os

import numpy as np
import torch
import torch.distributed as dist

from . import dist as sdist
from .configuration import defaultConfig
from .training import TrainEngine
from ._lib import ScdError


class NetworkFactory(object):

    def __init__(self, useGPU=True, config=None, dataset=None):
        if not useGPU or not torch.cuda.is_available():
            raise ScdError("NetworkFactory (scd_b200) runs on a B200 only: there is no CPU path")
        self.config = config or defaultConfig
        self.useGPU = True
        plugin = importlib.import_module(self.config.dirModel)           # ref: networkFactory.py:50-57
        self.model = plugin.model(**plugin.modelParams)
        self.loss = plugin.loss
        self.evaluation = getattr(plugin, "evaluation", None)
        self.evalExpr = getattr(plugin, "expression", None)
        self.dataset = dataset
        self.parameterCount = sum(p.numel() for p in self.model.parameters())   # ref: :70-77
        if self.config.optimizer != "adam":
            raise ScdError("only the reference's default optimizer (adam) is built")
        self.engine = None
        self.learningRate = None

    # ------------------------------------------------------------------ set-up (ref: :126-144)
    def prepare(self, localRank=0):
        dev = torch.device("cuda", localRank if localRank >= 0 else 0)
        torch.cuda.set_device(dev)
        self.model = self.model.to(dev)
        if self.config.pretrain is not None:
            self.loadPretrained(os.path.join(self.config.dirPretrain, self.config.pretrain))
        sdist.broadcast_module(self.model, 0)
        self.model.train()
        group = dist.group.WORLD if dist.is_initialized() and dist.get_world_size() > 1 else None
        # quirk kept from the reference: Adam starts from torch's default lr 1e-3, the configured learningRate
        # only takes effect at the first decay (networkFactory.py:80-82 vs :228-231)
        self.engine = TrainEngine(self.model, lr=1e-3, regr_w=self.loss.regressionWeight,
                                  off_w=self.loss.offsetWeight, process_group=group)
        self.learningRate = self.config.learningRate
        return self

    # ------------------------------------------------------------------ the loop (ref: :99-241)
    def beginTraining(self, localRank=0, on_iteration=None):
        if self.engine is None:
            self.prepare(localRank)
        cfg = self.config
        it = cfg.currentIter
        decay_at, decay_rate = list(cfg.learningRateDecay), list(cfg.learningRateDecayRate)
        log = []
        finished = it >= cfg.iterations
        while not finished:
            for data in self.dataset:
                cfg.updateIteration(it)
                it += 1
                loss, stats = self.train(**data)
                log.append([it, loss] + stats)
                if on_iteration is not None:
                    on_iteration(it, loss, stats)
                if it % cfg.snapshot == 0:
                    self.saveParameters()
                    arr = np.asarray([[r[0]] + [float(v) for v in r[1:]] for r in log], np.float64)
                    np.savetxt(os.path.join(cfg.directory("dirResult"), "losses.%s.%d.txt" % (cfg.trainName, it)),
                               arr, delimiter=",", fmt="%.5f")
                    log = []
                if decay_at and it == decay_at[0]:                      # ref: :225-234
                    self.learningRate /= decay_rate[0]
                    self.setLearningRate(self.learningRate)
                    decay_at.pop(0)
                    decay_rate.pop(0)
                if it >= cfg.iterations:
                    finished = True
                    break
        return it

    def train(self, xs, ys, **kwargs):
        """ref: NetworkFactory.train :257-263.  Returns (loss 0-dim tensor, [focal, size, offset]) on the device."""
        losses = self.engine.train_step(xs[0], ys)
        return losses[0], [losses[1], losses[2], losses[3]]

    def validate(self, xs, ys, **kwargs):
        """ref: :265-271: decode in eval-free no_grad mode.  The reference validates with the module in train mode
        (batch statistics); here validation uses the running statistics (eval mode), the mode inference runs in."""
        was = self.model.training
        self.model.eval()
        try:
            with torch.no_grad():
                result = self.model(*xs, decode=True)
        finally:
            self.model.train(was)
        if self.evaluation is None or ys is None:
            return result
        return self.evaluation(xs, ys, *result)

    def setLearningRate(self, lr):
        self.engine.set_learning_rate(lr)

    # ------------------------------------------------------------------ checkpoints (ref: :273-302)
    def _state_dict(self):
        """Keys carry the `module.` prefix of the reference's DDP-wrapped checkpoints (trace.py -wrapped)."""
        return {"module." + k: v.detach().clone() for k, v in self.model.state_dict().items()}

    def saveParameters(self):
        path = os.path.join(self.config.directory("dirTemp"), self.config.naming)
        if not dist.is_initialized() or dist.get_rank() == 0:
            torch.save(self._state_dict(), path)
        return path

    def _load(self, path):
        sd = torch.load(path, map_location="cpu")
        sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in sd.items()}
        with torch.no_grad():
            for k, v in self.model.state_dict().items():
                v.copy_(sd[k])                         # in place: parameters may be views into the engine's buffer
        if self.engine is not None:
            self.engine.refresh_operands()

    def loadParameters(self):
        self._load(os.path.join(self.config.dirTemp, self.config.naming))

    def loadPretrained(self, pretrained):
        self._load(pretrained)
