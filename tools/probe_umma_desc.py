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


if __name__ == "__main__" and "--mn" not in sys.argv:
    main()


def mn_major_probe():
    """MN-major operands with the 128-byte swizzle read as WINDOWS of a pixel strip (wgrad strip mode): the operand start
    is a whole number of 128 B pixel rows into a strip written the way TMA writes it (16 B chunk c of row r at chunk
    c ^ (r & 7), keyed on the ADDRESS bits), the two 64-channel halves of the M = 128 tile are two different windows
    (LBO = their distance), the K = 16 pixels are two 8-row groups 1024 B apart."""
    rng = np.random.default_rng(1)
    npx = 160                                               # pixel rows of 64 channels (128 B)
    vals = rng.integers(-4, 5, size=(2, npx, 64)).astype(np.float16)          # [A strip, B strip][pixel][channel]
    img16 = np.zeros(IMG // 2, np.float16)
    for which, base in ((0, A_OFF), (1, B_OFF)):
        for r in range(npx):
            for c in range(8):
                off = base + r * 128 + ((c ^ (r & 7)) << 4)
                img16[off // 2: off // 2 + 8] = vals[which, r, c * 8:(c + 1) * 8]
    res = []
    for p0, p1, q0 in ((0, 8, 0), (1, 19, 0), (3, 22, 5), (18, 37, 2), (37, 38, 9)):
        a_start, lbo = A_OFF + p0 * 128, (p1 - p0) * 128
        ad = (a_start >> 4) | ((lbo >> 4) << 16) | ((1024 >> 4) << 32) | (1 << 46) | (2 << 61)
        b_start = B_OFF + q0 * 128
        bd = (b_start >> 4) | ((8192 >> 4) << 16) | ((1024 >> 4) << 32) | (1 << 46) | (2 << 61)
        idesc_mn = idesc(128, 64) | (1 << 15) | (1 << 16)
        img = torch.from_numpy(img16.view(np.uint8)).cuda()
        out = torch.empty(128, 64, dtype=torch.float32, device="cuda")
        S.check(S.lib.scd_probe_umma(ctypes.c_void_p(img.data_ptr()), img.numel(), ad, bd, idesc_mn, 64, 1, 0, 0,
                                     ctypes.c_void_p(out.data_ptr()),
                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "scd_probe_umma")
        torch.cuda.synchronize()
        got = out.cpu().numpy().astype(np.float64)
        A = np.concatenate([vals[0, p0:p0 + 16].astype(np.float64), vals[0, p1:p1 + 16].astype(np.float64)], axis=1)   # (16 px, 128 ch)
        B = vals[1, q0:q0 + 16].astype(np.float64)                                                                       # (16 px, 64 ch)
        exp = A.T @ B
        row = {"case": "MN-major window", "p0": p0, "p1": p1, "q0": q0, "match": bool(np.array_equal(got, exp)),
               "max_abs_diff": float(np.abs(got - exp).max())}
        res.append(row)
        print(json.dumps(row))
    return res


if __name__ == "__main__" and "--mn" in sys.argv:
    mn_major_probe()
