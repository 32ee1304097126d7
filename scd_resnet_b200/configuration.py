"""Run configuration with the keys of the reference's JSON files (ref: configuration.py:9-44, configs/exp74.json).

Only keys that already exist are updated from a JSON object, like the reference's updateConfig
(configuration.py:150-153); templated entries ({modelName}, {trainName} ...) are expanded on read."""
import json
import os

_DEFAULTS = {
    "datasetName": None, "modelName": "centerOffsetRes10", "trainName": "run",
    "learningRate": 0.00025, "learningRateDecay": [80000], "learningRateDecayRate": [10],
    "currentIter": 0, "iterations": 117000, "validation": 200, "snapshot": 2000,
    "batchSize": 32, "validationBatchSize": 160,
    "naming": "{modelName}.{trainName}.{currentIter}.pth", "namingOptimizer": "{naming}.{optimizer}.pth",
    "pretrain": None, "optimizer": "adam",
    "dirData": "trainer.dataset.{datasetName}", "dirModel": "scd_resnet_b200.trainer.model.{modelName}",
    "dirTemp": "/tmp/scd_b200/temp/", "dirPretrain": "/tmp/scd_b200/pretrain/", "dirResult": "/tmp/scd_b200/results/",
    "dirConfig": "/tmp/scd_b200/configs/", "dirDataset": "/tmp/scd_b200/datasets/",
    "dirDatafile": "{dirDataset}{datasetName}.d", "dirDataSplitProfile": "{dirDataset}{datasetName}.split.json",
    "useGPU": True,
}


class Configuration:
    def __init__(self, **overrides):
        self.config = dict(_DEFAULTS)
        self.update_from(overrides)

    def update_from(self, obj):
        for k, v in obj.items():
            if k in self.config:
                self.config[k] = v
        return self

    updateConfig = update_from                      # the reference's method name

    def load(self, path):
        with open(path) as f:
            return self.update_from(json.load(f))

    def expand(self, key):
        v = self.config[key]
        for _ in range(3):                          # templates may nest one level ("{naming}.{optimizer}.pth")
            if not isinstance(v, str) or "{" not in v:
                break
            v = v.format(**self.config)
        return v

    def __getattr__(self, key):
        cfg = self.__dict__.get("config", {})
        if key in cfg:
            return self.expand(key)
        raise AttributeError(key)

    def directory(self, key):
        d = self.expand(key)
        os.makedirs(d, exist_ok=True)
        return d

    def updateIteration(self, it):
        self.config["currentIter"] = it


defaultConfig = Configuration()
