"""Which association order do CUDA-eager `torch.sum(w, -1)` and `torch.cumsum(pdf, -1)` use on rows of 64 / 128 floats
(neural_rendering.py:189-191, the importance-sampling cdf)?  Candidate orders are evaluated with plain fp32 adds on the
CPU and compared bit for bit with the device results.  python scripts/cdf_probe.py [out.json]"""
import json, sys
import torch

dev = torch.device("cuda", 0)


def eq_frac(a, b):
    return float((a.contiguous().view(torch.int32) == b.contiguous().view(torch.int32)).float().mean())


# ------------------------------------------------------------------ row sums
def sum_seq(w):
    s = w[:, 0].clone()
    for k in range(1, w.shape[1]):
        s = s + w[:, k]
    return s


def tree(t, ascending):
    """lane i += lane i + off (shfl_down); lanes past the end keep their value; lane 0 is the result"""
    n = t.shape[1]
    offs = [1 << i for i in range((n - 1).bit_length())]
    for off in (offs if ascending else offs[::-1]):
        nxt = t.clone()
        nxt[:, :n - off] = t[:, :n - off] + t[:, off:]
        t = nxt
    return t[:, 0]


def sum_vec(w, vec, ascending):
    v = w.view(w.shape[0], -1, vec)
    t = v[:, :, 0].clone()
    for i in range(1, vec):
        t = t + v[:, :, i]
    return tree(t, ascending)


def sum_strided(w, nthreads, ascending):
    v = w.view(w.shape[0], -1, nthreads)           # element k -> thread k % nthreads
    t = v[:, 0, :].clone()
    for i in range(1, v.shape[1]):
        t = t + v[:, i, :]
    return tree(t, ascending)


def sum_strided_acc4(w, nthreads, ascending):
    """thread-strided with up to 4 independent accumulators per thread (vt0 = 4), combined in order"""
    v = w.view(w.shape[0], -1, nthreads)
    accs = [None] * 4
    for i in range(v.shape[1]):
        a = i % 4
        accs[a] = v[:, i, :].clone() if accs[a] is None else accs[a] + v[:, i, :]
    t = accs[0]
    for a in accs[1:]:
        if a is not None:
            t = t + a
    return tree(t, ascending)


# ------------------------------------------------------------------ row scans
def scan_seq(p):
    out = p.clone()
    for k in range(1, p.shape[1]):
        out[:, k] = out[:, k - 1] + p[:, k]
    return out


def scan_chunks(p, ntx, kind):
    """ATen's tensor_kernel_scan_innermost_dim: chunks of 2 * ntx elements, the running total added to the chunk's first
    element, then an in-chunk network: 'sklansky' or 'blelloch' (up-sweep / down-sweep)"""
    R, K = p.shape
    out = torch.empty_like(p)
    total = torch.zeros(R)
    W = 2 * ntx
    for c0 in range(0, K, W):
        buf = torch.zeros(R, W)
        n = min(W, K - c0)
        buf[:, :n] = p[:, c0:c0 + n]
        buf[:, 0] = buf[:, 0] + total if c0 > 0 else buf[:, 0]
        if kind == "sklansky":
            s = 1
            while s <= ntx:
                new = buf.clone()
                for t in range(ntx):
                    a = (t // s) * (2 * s) + s
                    ti, si = a + (t % s), a - 1
                    new[:, ti] = buf[:, ti] + buf[:, si]
                buf = new
                s <<= 1
        else:
            s, d = ntx, 1
            while s >= 1:
                new = buf.clone()
                for t in range(s):
                    off = (2 * t + 1) * d - 1
                    new[:, off + d] = buf[:, off] + buf[:, off + d]
                buf = new
                s >>= 1
                d <<= 1
            s, d = 2, ntx // 2
            while d >= 1:
                new = buf.clone()
                for t in range(s - 1):
                    off = 2 * (t + 1) * d - 1
                    new[:, off + d] = buf[:, off] + buf[:, off + d]
                buf = new
                s <<= 1
                d >>= 1
        out[:, c0:c0 + n] = buf[:, :n]
        total = buf[:, W - 1].clone()
    return out


res = {}
for Kc in (64, 128):
    g = torch.Generator().manual_seed(Kc)
    R = 4096
    # compositing weights: alpha * transmittance of random densities, + 1e-5 (what sample_fine sees)
    sig = torch.rand(R, Kc, generator=g) * 3 * (torch.rand(R, Kc, generator=g) < 0.5)
    alpha = 1 - torch.exp(-sig * 0.03)
    T = torch.cumprod(torch.cat([torch.ones(R, 1), 1 - alpha + 1e-10], 1), 1)[:, :-1]
    w = (alpha * T + 1e-5).float()
    wd = w.to(dev)
    s_ref = torch.sum(wd, -1).cpu()
    sums = {"sequential": sum_seq(w)}
    for asc in (True, False):
        tag = "asc" if asc else "desc"
        sums[f"vec4 + tree {tag}"] = sum_vec(w, 4, asc)
        sums[f"vec2 + tree {tag}"] = sum_vec(w, 2, asc)
        for nt in (16, 32, 64):
            if Kc % nt == 0 and nt <= Kc:
                sums[f"strided over {nt} threads + tree {tag}"] = sum_strided(w, nt, asc)
                sums[f"strided over {nt} threads, 4 accumulators + tree {tag}"] = sum_strided_acc4(w, nt, asc)
    res[f"sum Kc={Kc}"] = {k: eq_frac(v, s_ref) for k, v in sums.items()}
    pdf_d = wd / torch.sum(wd, -1, keepdim=True)
    pdf = pdf_d.cpu()                                   # the device's own pdf: isolates the scan
    c_ref = torch.cumsum(pdf_d, -1).cpu()
    scans = {"sequential": scan_seq(pdf)}
    for ntx in (16, 32, 64, 128, 256, 512):
        for kind in ("sklansky", "blelloch"):
            scans[f"{kind}, {ntx} threads"] = scan_chunks(pdf, ntx, kind)
    res[f"cumsum Kc={Kc}"] = {k: eq_frac(v, c_ref) for k, v in scans.items()}
    # does the division agree with an IEEE division of the same operands?
    res[f"pdf Kc={Kc} (w / device sum, IEEE division on the CPU)"] = eq_frac(w / s_ref[:, None], pdf)
for k, v in res.items():
    if isinstance(v, dict):
        best = max(v, key=v.get)
        print(f"{k}: best '{best}' = {v[best]:.6f}; exact: {[n for n, f in v.items() if f == 1.0]}")
    else:
        print(f"{k}: {v:.6f}")
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
