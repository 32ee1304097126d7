"""Run-to-run determinism of the inference path (GPU box only).

The detector has no floating-point atomics on the forward path, so the same tiles must give bit-identical maps and
planes every time.  This script repeats (a) detect_device on fixed device tiles, (b) detect_host on fixed pinned tiles
(float32 and uint8, pipelined copies), comparing every repetition with the first, and prints the mismatch counts.
Environment toggles read by the library (SCD_PDL, SCD_STEM_IMPL, SCD_IGEMM_ROW_MODE) localise a difference."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import scd_resnet_b200 as S
from scd_resnet_b200 import synthetic
from scd_resnet_b200.centerNetOffset import CenterNetResidual
from scd_resnet_b200.inference import TileDetector


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    m = CenterNetResidual(10)
    m.load_state_dict(synthetic.make_state_dict(m, 1234))
    m.cuda().eval()
    det = TileDetector(m, batch, "cuda")
    rng = np.random.default_rng(17)
    u8 = [torch.from_numpy(rng.integers(0, 256, size=(batch, 1, 512, 512), dtype=np.uint8)).pin_memory() for _ in range(3)]
    x = torch.stack([torch.from_numpy(rng.standard_normal((1, 512, 512)).astype(np.float32)) for _ in range(batch)]).cuda()
    f32 = [torch.randn(batch, 1, 512, 512).pin_memory() for _ in range(3)]
    out = {"reps": reps, "batch": batch, "env": {k: os.environ.get(k) for k in ("SCD_PDL", "SCD_STEM_IMPL", "SCD_IGEMM_ROW_MODE")}}

    ref_maps, bad_maps, bad_planes = None, 0, 0
    for _ in range(reps):
        planes = det.detect_device(x).clone()
        maps = [t[:batch].clone() for t in det.maps]
        torch.cuda.synchronize()
        if ref_maps is None:
            ref_maps, ref_planes = maps, planes
        else:
            bad_maps += int(any(not torch.equal(a, b) for a, b in zip(maps, ref_maps)))
            bad_planes += int(not torch.equal(planes, ref_planes))
    out["detect_device"] = {"maps_differ": bad_maps, "planes_differ": bad_planes}

    for name, batches in (("detect_host_u8", u8), ("detect_host_f32", f32)):
        ref, bad = None, 0
        for _ in range(reps // 4):
            got = [p.clone() for p in det.detect_host(batches)]
            if ref is None:
                ref = got
            else:
                bad += int(any(not torch.equal(a, b) for a, b in zip(got, ref)))
        out[name] = {"runs": reps // 4, "differ": bad}
    # three batches in flight, repeated: planes AND the whole activation workspace of the last batch
    ref, bad = None, 0
    for _ in range(reps // 2):
        got = [p.clone() for p in det.detect_host(u8)]
        torch.cuda.synchronize()
        ws = det.workspace.clone()
        if ref is None:
            ref = (got, ws)
        else:
            bad += int(not torch.equal(ws, ref[1]) or any(not torch.equal(a, b) for a, b in zip(got, ref[0])))
    out["detect_host_u8_workspace"] = {"runs": reps // 2, "differ": bad}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
