"""Measures every distinct GEMM / conv of the UNet plans (B_eff = 2 and 16 by default) under each
(tile width, K-splits, ring depth) candidate and writes the winners to
diff_foley_b200/csrc/igemm_tuned.inc, the planner's measured table.

Each candidate is timed as one CUDA-graph replay of 12 launches over 12 different weight copies
(the UNet streams every weight exactly once per forward, so operands must come from HBM, not L2),
with an L2 flush before each timed replay; best of 3.
    python tools/autotune_igemm.py [b_eff ...]        (needs a B200; rebuild the library afterwards)"""
import sys

sys.path.insert(0, ".")
import torch

from bench import FULL
from diff_foley_b200 import _lib as L
from diff_foley_b200.unet import UNetModelB200
from diff_foley_b200.weights import randomize_parameters_

dev = torch.device("cuda", 0)
lib = L.lib()
NREP = 12
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
LEVELS = {1024: (16, 64), 256: (8, 32), 64: (4, 16), 16: (2, 8)}


def shapes_for(b_eff):
    unet = UNetModelB200(**FULL, max_batch=max(b_eff, 2)).to(dev)
    randomize_parameters_(unet, 7)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(b_eff, 4, 16, 64, generator=g).to(dev)
    ctx = torch.randn(b_eff, 32, 768, generator=g).to(dev)
    t = torch.full((b_eff,), 961, device=dev, dtype=torch.long)
    lib.dfb_debug_igemm_force(0, -1)
    unet(x, t, context=ctx)
    prof = unet.profile(x, t, ctx, iters=1)
    out = {}
    for i, p in enumerate(prof):
        if not p["kind"].startswith("igemm"):
            continue
        taps = 9 if p["kind"] == "igemm_conv3x3" else 1
        # GEGLU launches are the linears whose N is 8C with K = C
        geglu = int(taps == 1 and p["N"] == 8 * p["K"])
        key = (p["M"], p["N"], p["K"], taps, geglu)
        d = out.setdefault(key, dict(count=0, cur=(p["splits"], p["ctas"])))
        d["count"] += 1
    unet.release()
    del unet
    torch.cuda.empty_cache()
    return out


def time_candidate(key, b_eff, bn, splits, deep):
    M, N, K, taps, geglu = key
    lib.dfb_debug_igemm_force(bn, deep)
    nw = NREP
    ws = [(torch.randn(N, K, device=dev) / K ** 0.5).half() for _ in range(nw)]
    C2 = 0
    if taps == 9:
        # conv2 of a ResBlock with a fused 1x1 skip connection: K = 9*N + C2 (never a multiple of 9)
        C, C2 = (K // 9, 0) if K % 9 == 0 else (N, K - 9 * N)
        H, W = LEVELS[M // b_eff]
        a = torch.randn(b_eff, H, W, C, device=dev).half()
        a2 = torch.randn(b_eff, H, W, C2, device=dev).half() if C2 else None
    else:
        a = torch.randn(M, K, device=dev).half()
    No = N // 2 if geglu else N
    bias = torch.randn(N, device=dev)
    res = None if geglu else torch.randn(M, No, device=dev)
    o32 = None if geglu else torch.empty(M, No, device=dev)
    o16 = torch.empty(M, No, device=dev, dtype=torch.float16) if geglu else None

    def launch(w):
        if taps == 9 and C2:
            return lib.dfb_conv3x3_cat(L.ptr(a), L.ptr(a2), C2, L.ptr(w), b_eff, H, W, C, N, L.ptr(bias), None,
                                       None, L.ptr(o32), None, splits, L.cur_stream())
        if taps == 9:
            return lib.dfb_conv3x3(L.ptr(a), L.ptr(w), b_eff, H, W, C, N, L.ptr(bias), None, L.ptr(res), 0,
                                   L.ptr(o32), None, splits, L.cur_stream())
        return lib.dfb_gemm(L.ptr(a), L.ptr(w), M, N, K, L.ptr(bias), L.ptr(res), 2 if geglu else 0,
                            L.ptr(o32), L.ptr(o16), splits, L.cur_stream())

    if launch(ws[0]) != 0:
        return None
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for w in ws:
            L.check(launch(w))
    best = 1e9
    for _ in range(3):
        flush_buf.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / nw)
    return best


def main():
    b_effs = [int(a) for a in sys.argv[1:]] or [2, 16]
    rows = []
    total_cur = total_best = 0.0
    for b_eff in b_effs:
        shapes = shapes_for(b_eff)
        print(f"# b_eff={b_eff}: {len(shapes)} distinct igemm shapes", flush=True)
        for key, info in sorted(shapes.items()):
            M, N, K, taps, geglu = key
            kb_total = K // 64
            res = {}
            cur_t = time_candidate(key, b_eff, 0, 0, -1)  # the cost model's own choice
            for bn in ((128,) if geglu else (64, 128)):
                if bn == 128 and N <= 64:
                    continue
                for sp in range(1, 9):
                    if sp > kb_total:
                        break
                    for deep in (0, 1):
                        if bn == 128 and sp > 1 and deep == 0:
                            continue  # needs the deep ring's shared memory
                        t = time_candidate(key, b_eff, bn, sp, deep)
                        if t is not None:
                            res[(bn, sp, deep)] = t
            if not geglu and M >= 1024 and N >= 144:   # wide tiles (128 x up-to-256 columns, 1 CTA / SM, no split-K)
                t = time_candidate(key, b_eff, 256, 1, 1)
                if t is not None:
                    res[(256, 1, 1)] = t
            (bn, sp, deep), t = min(res.items(), key=lambda kv: kv[1])
            total_cur += cur_t * info["count"]
            total_best += min(t, cur_t) * info["count"]
            tag = "" if t < 0.97 * cur_t else "  (kept: cost model)"
            print(f"M={M:6d} N={N:6d} K={K:6d} taps={taps} geglu={geglu} x{info['count']:2d}: model {cur_t:7.2f} us "
                  f"(splits {info['cur'][0]}, {info['cur'][1]} CTAs) | best {t:7.2f} us BN={bn} splits={sp} deep={deep}{tag}", flush=True)
            if t < 0.97 * cur_t:
                rows.append((M, N, K, taps, geglu, bn, sp, deep, cur_t, t))
    lib.dfb_debug_igemm_force(0, -1)
    print(f"# sum over launches: cost model {total_cur:.0f} us -> tuned {total_best:.0f} us")
    path = "gpurun_out/igemm_tuned.inc" if len(sys.argv) <= 1 or True else None
    with open(path, "w") as f:
        f.write("// generated by tools/autotune_igemm.py -- {M, N, K, taps, geglu, BN, splits, deep}\n")
        for r in rows:
            f.write("    {%d, %d, %d, %d, %d, %d, %d, %d},  // %.2f -> %.2f us\n" % r)
    print("wrote", path, "- copy it to diff_foley_b200/csrc/igemm_tuned.inc and rebuild")


if __name__ == "__main__":
    main()
