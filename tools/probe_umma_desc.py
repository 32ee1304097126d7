"""How does tcgen05.mma read a K-major operand WITHOUT swizzle?  (GPU box only.)

The canonical layout (cute/atom/mma_traits_sm100.hpp) is ((8,m),(T,2)):((1T,SBO),(1,LBO)) in 16-byte units: the 8 rows
of a core matrix sit 16 B apart, K-adjacent core matrices LBO apart, M-adjacent 8-row groups SBO apart.  The stem
kernel wants to read its A operand IN PLACE from the space-to-depth patch, which needs: arbitrary (16 B granular) start
addresses, LBO = 16 B (core matrices that overlap: chunk 1 of row i is chunk 0 of row i + 1), SBO = 128 B or a row
pitch, and a B operand in a different layout type (128-byte swizzle) in the same instruction.  This script runs the
probe kernel (scd_probe_umma) on random integer-valued fp16 data for each of those and prints which reading of
(LBO, SBO) reproduces the result: H1 = LBO along K / SBO along M (the CUTLASS comment), H2 = swapped."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import scd_resnet_b200 as S

A_OFF, B_OFF, IMG = 0, 65536, 98304


def desc(start, lbo, sbo, layout=0):
    return (start >> 4) | ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1 << 46) | (layout << 61)


def idesc(m, n, afmt=0, bfmt=0):
    return (1 << 4) | (afmt << 7) | (bfmt << 10) | ((n >> 3) << 17) | ((m >> 4) << 24)


def read_nosw(img16, start, rows, lbo, sbo, k0=0, k=16):
    """(rows, k) matrix read under H1: element (r, kk) at start + (r%8)*16 + (r//8)*sbo + (kk//8)*lbo + (kk%8)*2."""
    r = np.arange(rows)[:, None]
    kk = np.arange(k)[None, :]
    off = start + (r % 8) * 16 + (r // 8) * sbo + ((kk + k0) // 8) * lbo + ((kk + k0) % 8) * 2
    return img16[off // 2].astype(np.float64)


def read_sw128(img16, start, rows, k0=0, k=16):
    """(rows, k) from a 128-byte-swizzled K-major tile (rows of 128 B, chunk j of row n at ((j ^ (n & 7)) << 4))."""
    r = np.arange(rows)[:, None]
    kk = np.arange(k)[None, :] + k0
    off = start + r * 128 + (((kk // 8) ^ (r & 7)) << 4) + (kk % 8) * 2
    return img16[off // 2].astype(np.float64)


def run(img16, adesc, bdesc, n=64, k_steps=1, a_step=0, b_step=0):
    img = torch.from_numpy(img16.view(np.uint8)).cuda()
    out = torch.empty(128, n, dtype=torch.float32, device="cuda")
    S.check(S.lib.scd_probe_umma(ctypes.c_void_p(img.data_ptr()), img.numel(), adesc, bdesc, idesc(128, n), n, k_steps,
                                 a_step, b_step, ctypes.c_void_p(out.data_ptr()),
                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "scd_probe_umma")
    torch.cuda.synchronize()
    return out.cpu().numpy().astype(np.float64)


def main():
    rng = np.random.default_rng(0)
    img16 = rng.integers(-4, 5, size=IMG // 2).astype(np.float16)
    res = []

    def case(name, a_start, a_lbo, a_sbo, b_kind="nosw", k_steps=1, a_step_bytes=0):
        if b_kind == "nosw":
            bd = desc(B_OFF, 128, 256)
            Bm = [read_nosw(img16, B_OFF, 64, 128, 256)] * k_steps
            b_step = 0
        else:
            bd = desc(B_OFF, 0, 1024, 2)
            Bm = [read_sw128(img16, B_OFF, 64, 16 * s) for s in range(k_steps)]
            b_step = 2
        ad = desc(A_OFF + a_start, a_lbo, a_sbo)
        got = run(img16, ad, bd, 64, k_steps, a_step_bytes >> 4, b_step)
        h1 = sum(read_nosw(img16, A_OFF + a_start + s * a_step_bytes, 128, a_lbo, a_sbo) @ Bm[s].T for s in range(k_steps))
        h2 = sum(read_nosw(img16, A_OFF + a_start + s * a_step_bytes, 128, a_sbo, a_lbo) @ Bm[s].T for s in range(k_steps))
        row = {"case": name, "a_start": a_start, "a_lbo": a_lbo, "a_sbo": a_sbo, "b": b_kind, "k_steps": k_steps,
               "H1_lbo_is_K": bool(np.array_equal(got, h1)), "H2_swapped": bool(np.array_equal(got, h2)),
               "max_abs_diff_H1": float(np.abs(got - h1).max())}
        res.append(row)
        print(json.dumps(row))

    case("packed core matrices", 0, 128, 256)
    case("packed, start + 16 B", 16, 128, 256)
    case("packed, start + 1040 B", 1040, 128, 256)
    case("K chunks 32 B apart, groups 160 B apart", 0, 32, 160)
    case("rows contiguous: LBO 16 (overlapping), SBO 128", 0, 16, 128)
    case("the same, start + 48 B", 48, 16, 128)
    case("LBO 16, SBO = 2080 B row pitch", 0, 16, 2080)
    case("stem layout vs 128 B-swizzled B", 32, 16, 128, b_kind="sw128")
    case("stem chain: 4 K steps, A advances one 2080 B patch row per step, B swizzled", 32, 16, 128, b_kind="sw128",
         k_steps=4, a_step_bytes=2080)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "probe_umma_desc.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
