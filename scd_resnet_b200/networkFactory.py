"""Training loop with the surface of the reference's NetworkFactory (ref: models/networkFactory.py:36-302):
plugin lookup by module path, train / validate, periodic validation with the `evals.{trainName}.txt` report,
learning-rate decay, snapshots, resume from `currentIter`.

Two ways to run the step, same kernels underneath:

* engine="native" (default): `train` is one TrainEngine step (forward with batch-statistics BatchNorm, fused
  CenterNetLoss, backward, fused Adam over the flat parameter buffer).  With several ranks the engine averages the flat
  gradient buffer over NCCL and shares the BatchNorm statistics: the DDP + SyncBatchNorm semantics of
  networkFactory.py:126-134.  Parameters are broadcast from rank 0 at start, as DDP does.
* engine="autograd": the reference's own sequence, line for line (networkFactory.py:126-134, 257-263): the module is
  converted with torch.nn.SyncBatchNorm.convert_sync_batchnorm, wrapped in DistributedDataParallel, and
  `zero_grad -> model(*xs, decode=False) -> loss -> loss.mean().backward() -> optimizer.step()` runs with a torch
  optimizer (Adam or SGD) on the module's parameters.  The module's train-mode forward is one autograd node whose
  backward is the native backward pass (centerNetOffset._TrainForwardFn).  This is the route an unmodified reference
  NetworkFactory takes when `dirModel` points at this package's plugin.
"""
import importlib
import json
import os

import numpy as np
import torch
import torch.distributed as dist

from . import dist as sdist
from .configuration import defaultConfig
from .training import TrainEngine
from ._lib import ScdError


class NetworkFactory(object):

    def __init__(self, useGPU=True, config=None, dataset=None, engine="native"):
        if not useGPU or not torch.cuda.is_available():
            raise ScdError("NetworkFactory (scd_b200) runs on a B200 only: there is no CPU path")
        if engine not in ("native", "autograd"):
            raise ScdError("engine must be 'native' or 'autograd'")
        self.config = config or defaultConfig
        self.useGPU = True
        self.mode = engine
        plugin = importlib.import_module(self.config.dirModel)           # ref: networkFactory.py:50-57
        self.model = plugin.model(**plugin.modelParams)
        self.loss = plugin.loss
        self.evaluation = getattr(plugin, "evaluation", None)
        self.evalExpr = getattr(plugin, "expression", None)
        self.dataset = dataset if dataset is not None else self._load_dataset()      # ref: :59-68
        self.parameterCount = sum(p.numel() for p in self.model.parameters())   # ref: :70-77
        self.optimizer = None
        if engine == "autograd":                                         # ref: :79-93
            params = [p for p in self.model.parameters() if p.requires_grad]
            if self.config.optimizer == "adam":
                self.optimizer = torch.optim.Adam(params)
            elif self.config.optimizer == "sgd":
                self.optimizer = torch.optim.SGD(params, lr=self.config.learningRate, momentum=0.9, weight_decay=0.0001)
            else:
                raise ScdError("Unknown Optimizer '%s', Currently Support 'sgd' or 'adam'" % self.config.optimizer)
        elif self.config.optimizer != "adam":
            raise ScdError("engine='native' builds the reference's default optimizer (adam); use engine='autograd' for sgd")
        self.engine = None
        self.learningRate = None
        self._prepared = False

    def _load_dataset(self):
        """ref: :59-68: the dataset plugin `dirData` exports dataset(dirDatafile, useGPU, splitProfile).  Without a
        datasetName there is nothing to load and the caller passes a dataset object instead."""
        cfg = self.config
        if cfg.datasetName is None:
            return None
        loader = importlib.import_module(cfg.dirData)
        split = None
        if os.path.exists(cfg.dirDataSplitProfile):
            with open(cfg.dirDataSplitProfile, "r") as f:
                split = json.load(f)
        return loader.dataset(cfg.dirDatafile, True, split)

    @property
    def isGPU(self):
        return self.useGPU

    # ------------------------------------------------------------------ set-up (ref: :113-146)
    def _resume_learning_rate(self):
        """Learning rate after the decay milestones that lie at or before `currentIter` (ref: :116-124; the reference
        indexes learningRateDecayRate with the iteration number there, which can only raise: the intent is applied)."""
        cfg = self.config
        lr = cfg.learningRate
        pending_at, pending_rate = [], []
        for at, rate in zip(cfg.learningRateDecay, cfg.learningRateDecayRate):
            if cfg.currentIter > 0 and at <= cfg.currentIter:
                lr /= rate
            else:
                pending_at.append(at)
                pending_rate.append(rate)
        return lr, pending_at, pending_rate

    def prepare(self, localRank=0):
        cfg = self.config
        dev = torch.device("cuda", localRank if localRank >= 0 else 0)
        torch.cuda.set_device(dev)
        self.learningRate, self._decay_at, self._decay_rate = self._resume_learning_rate()
        resumed = cfg.currentIter > 0
        multi = dist.is_initialized() and dist.get_world_size() > 1
        if self.mode == "autograd":
            if resumed:                                                   # ref: :117-124 (before .cuda(), like there)
                self.loadParameters()
                self.setLearningRate(self.learningRate)
            self.model = self.model.to(dev)                               # ref: :126-134
            if dist.is_initialized():
                if multi:
                    self.model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(self.model)
                self.model = torch.nn.parallel.DistributedDataParallel(self.model, device_ids=[dev.index],
                                                                       find_unused_parameters=True)
            if cfg.pretrain is not None:
                self.loadPretrained(os.path.join(cfg.dirPretrain, cfg.pretrain))
            self.model.train()
        else:
            self.model = self.model.to(dev)
            if resumed:
                self.loadParameters()
            elif cfg.pretrain is not None:
                self.loadPretrained(os.path.join(cfg.dirPretrain, cfg.pretrain))
            sdist.broadcast_module(self.model, 0)
            self.model.train()
            group = dist.group.WORLD if multi else None
            # quirk kept from the reference: Adam starts from torch's default lr 1e-3, the configured learningRate
            # only takes effect at the first decay or on resume (networkFactory.py:80-82 vs :116-124, :228-231)
            self.engine = TrainEngine(self.model, lr=self.learningRate if resumed else 1e-3,
                                      regr_w=self.loss.regressionWeight, off_w=self.loss.offsetWeight,
                                      process_group=group)
        self._prepared = True
        return self

    @property
    def _net(self):
        """The plugin module underneath a DistributedDataParallel wrapper."""
        return self.model.module if hasattr(self.model, "module") else self.model

    # ------------------------------------------------------------------ the loop (ref: :148-241)
    def beginTraining(self, localRank=0, on_iteration=None):
        if not self._prepared:
            self.prepare(localRank)
        if self.dataset is None:
            raise ScdError("NetworkFactory.beginTraining: no dataset (set datasetName / dirData in the configuration or "
                           "pass dataset=...)")
        cfg = self.config
        it = cfg.currentIter
        decay_at, decay_rate = list(self._decay_at), list(self._decay_rate)
        log = []
        evalResult = ["Experiment: {}".format(cfg.trainName) + '\n',
                      "Parameter Count: {}".format(self.parameterCount) + '\n']      # ref: :156-157
        rank0 = not dist.is_initialized() or dist.get_rank() == 0
        finished = it >= cfg.iterations
        while not finished:
            for data in self.dataset:
                cfg.updateIteration(it)
                it += 1
                loss, stats = self.train(**data)
                log.append([it, loss] + list(stats))
                if on_iteration is not None:
                    on_iteration(it, loss, stats)
                if cfg.validation and it % cfg.validation == 0 and self.evaluation is not None:     # ref: :181-211
                    evalResult.append(self._validation_report(it, data))
                if it % cfg.snapshot == 0:                                                          # ref: :214-223
                    self.saveParameters()
                    if rank0:
                        arr = np.asarray([[r[0]] + [float(v) for v in r[1:]] for r in log], np.float64)
                        np.savetxt(os.path.join(cfg.directory("dirResult"), "losses.%s.%d.txt" % (cfg.trainName, it)),
                                   arr, delimiter=",", fmt="%.5f")
                    log = []
                    if dist.is_initialized():
                        dist.barrier()        # nobody runs ahead into a collective while rank 0 is still writing files
                if decay_at and it == decay_at[0]:                      # ref: :228-234
                    self.learningRate /= decay_rate[0]
                    self.setLearningRate(self.learningRate)
                    decay_at.pop(0)
                    decay_rate.pop(0)
                if it >= cfg.iterations:
                    finished = True
                    break
        if rank0:                                                                                   # ref: :240-241
            with open(os.path.join(cfg.directory("dirResult"), "evals.{}.txt".format(cfg.trainName)), "w") as f:
                f.writelines(evalResult)
        return it

    def _validation_report(self, it, data):
        """ref: :186-211: the current training batch and the validation split, both through validate() with the model
        left in train mode, each summarised by the plugin's `expression`."""
        trainResults, _ = self.validate(**data)
        evalTr = "[Tr] {}:     ".format(format(it, "7d")) + self.evalExpr([trainResults])
        batches = []
        with torch.no_grad():
            getter = getattr(self.dataset, "getValidationSet", None)
            for item in (getter() if getter is not None else []):
                results, _ = self.validate(**item)
                batches.append(results)
        evalr = "[It] {}:     ".format(format(it, "7d")) + (self.evalExpr(batches) if batches else "(no validation split)")
        return evalTr + '\n' + evalr + "\n"

    def cuda(self):
        self.model = self.model.cuda()

    def trainMode(self):
        self.model.train()

    def evalMode(self):
        self.model.eval()

    def _passParams(self, xs, ys, **kwargs):
        preds = self.model(*xs, **kwargs)
        loss, lossStats = self.loss(preds, ys)
        return loss, lossStats

    def train(self, xs, ys, **kwargs):
        """ref: NetworkFactory.train :257-263.  Returns (loss 0-dim tensor, [focal, size, offset]) on the device."""
        if self.mode == "autograd":
            self.optimizer.zero_grad()
            loss, lossStats = self._passParams(xs, ys, decode=False)
            loss = loss.mean()
            loss.backward()
            self.optimizer.step()
            return loss, lossStats
        losses = self.engine.train_step(xs[0], ys)
        return losses[0], [losses[1], losses[2], losses[3]]

    def validate(self, xs, ys, **kwargs):
        """ref: :265-271.  Like the reference, the module stays in whatever mode it is in: during training that is
        train mode, i.e. BatchNorm normalises with the statistics of the validation batch and moves its running
        statistics (the train-mode forward of the module handles no_grad without a tape)."""
        with torch.no_grad():
            result = self.model(*xs, **kwargs, decode=True)
        if self.evaluation is None or ys is None:
            return result
        return self.evaluation(xs, ys, *result)

    def setLearningRate(self, lr):
        if self.optimizer is not None:
            for group in self.optimizer.param_groups:
                group["lr"] = lr
        if self.engine is not None:
            self.engine.set_learning_rate(lr)

    # ------------------------------------------------------------------ checkpoints (ref: :278-302)
    def _state_dict(self):
        """Keys carry the `module.` prefix of the reference's DDP-wrapped checkpoints (trace.py -wrapped)."""
        return {"module." + k: v.detach().clone() for k, v in self._net.state_dict().items()}

    def saveParameters(self):
        path = os.path.join(self.config.directory("dirTemp"), self.config.naming)
        if not dist.is_initialized() or dist.get_rank() == 0:
            torch.save(self._state_dict(), path)
        return path

    def _load(self, path):
        sd = torch.load(path, map_location="cpu")
        sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in sd.items()}
        with torch.no_grad():
            for k, v in self._net.state_dict().items():
                v.copy_(sd[k])                         # in place: parameters may be views into the engine's buffer
        if self.engine is not None:
            self.engine.refresh_operands()

    def loadParameters(self):
        self._load(os.path.join(self.config.dirTemp, self.config.naming))

    def loadPretrained(self, pretrained):
        self._load(pretrained)
