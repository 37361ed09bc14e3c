"""First-contact diagnostics on a B200: every kernel against the CPU oracle, never stopping at
the first failure, so one gpurun call tells us as much as possible.

    python tools/gpu_diag.py [--only conv_tc]
"""
import argparse
import sys
import time
import traceback
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import vsr_oracle as O  # noqa: E402
from vsrlab_b200 import functional as VF  # noqa: E402
from vsrlab_b200 import ops  # noqa: E402
from vsrlab_b200._lib import ACT_LRELU, ACT_NONE, ACT_RELU, BF16, EPI_NHWC, F32  # noqa: E402

dev = torch.device("cuda:0")
RESULTS = []


def report(name, err, tol, extra=""):
    ok = err <= tol
    RESULTS.append((name, ok))
    print(f"[{'OK ' if ok else 'BAD'}] {name:58s} err={err:.3e} tol={tol:.1e} {extra}", flush=True)


def run(name, fn):
    try:
        fn()
    except Exception as e:  # noqa: BLE001
        RESULTS.append((name, False))
        print(f"[EXC] {name}: {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()
    st = ops.debug_status()
    if st:
        print(f"      !! pipeline time-out flag = {st} after {name}", flush=True)


def bf16r(x):
    return x.to(torch.bfloat16).to(torch.float32)


def conv_case(dt, cin_segs, cout, k, h, w, n=2, act="none", groups=1, pixshuf=0, residual=False, seed=0, tag=""):
    """segments given in OIHW order as [(off, c)]; input channels = sum(c)."""
    g = torch.Generator().manual_seed(seed)
    cin = sum(c for _, c in cin_segs)
    convs = []
    for _ in range(groups):
        cv = torch.nn.Conv2d(cin, cout, k, 1, k // 2)
        with torch.no_grad():
            cv.weight.copy_(torch.randn(cv.weight.shape, generator=g) * (1.0 / (cin * k * k) ** 0.5))
            cv.bias.copy_(torch.randn(cv.bias.shape, generator=g) * 0.1)
        convs.append(cv)
    B = n * groups
    x = torch.randn(B, cin, h, w, generator=g)
    res = torch.randn(B, cout, h, w, generator=g) if residual else None
    # oracle
    outs = []
    for gi in range(groups):
        xs = x[gi * n:(gi + 1) * n]
        wgt, b = convs[gi].weight.detach(), convs[gi].bias.detach()
        if dt == BF16:
            xs, wgt = bf16r(xs), bf16r(wgt)
        y = F.conv2d(xs, wgt, b, padding=k // 2)
        outs.append(y)
    y = torch.cat(outs)
    if act == "relu":
        y = y.clamp_min(0)
    elif act == "lrelu":
        y = torch.where(y >= 0, y, y * 0.1)
    if residual:
        y = y + (bf16r(res) if dt == BF16 else res)
    if pixshuf:
        y = O.pixel_shuffle(y, 2)
    # device
    tdt = ops.TORCH_DT[dt]
    gconvs = [c.to(dev) for c in convs]
    pc = ops.PackedConv(gconvs, cin_segs, dt, pixshuf)
    ins, in_c = [], []
    for off, c in cin_segs:
        ca = VF._act_c(c, dt)
        t = torch.zeros(B, h, w, ca, dtype=tdt, device=dev)
        t[..., :c] = x[:, off:off + c].permute(0, 2, 3, 1).to(dev).to(tdt)
        ins.append(t.contiguous())
        in_c.append(ca)
    r = pixshuf or 1
    co = cout // (r * r)
    oc = VF._act_c(co, dt) if r > 1 else pc.cout_pad
    out = torch.full((B, h * r, w * r, oc), 7.0, dtype=tdt, device=dev)
    rt = None
    rc = 0
    if residual:
        rc = VF._act_c(cout, dt)
        rt = torch.zeros(B, h, w, rc, dtype=tdt, device=dev)
        rt[..., :cout] = res.permute(0, 2, 3, 1).to(dev).to(tdt)
    ops.conv2d_fwd(pc, ins, in_c, B, h, w, act={"none": ACT_NONE, "relu": ACT_RELU, "lrelu": ACT_LRELU}[act], slope=0.1,
                   out=out, out_c=oc, residual=rt, res_c=rc)
    torch.cuda.synchronize()
    got = out[..., :co].float().permute(0, 3, 1, 2).cpu()
    err = (got - y).abs().max().item()
    scale = y.abs().max().item()
    tol = (2e-2 if dt == BF16 else 2e-5) * max(scale, 1.0)
    name = f"conv[{'bf16' if dt == BF16 else 'f32'}] k{k} segs{cin_segs}->{cout} {h}x{w} n{n} g{groups} {act}{' ps' if pixshuf else ''}{' res' if residual else ''} {tag}"
    report(name, err, tol, f"(max|y|={scale:.2f})")
    if err > tol and dt == BF16:
        d = (got - y).abs()
        bad = (d > tol)
        print("      bad fraction", bad.float().mean().item(), "per-channel max err", d.amax(dim=(0, 2, 3))[:8].tolist())
        print("      per-image max err", d.amax(dim=(1, 2, 3)).tolist())
        print("      err rows(y) max", d.amax(dim=(0, 1, 3))[:20].tolist())
        print("      err cols(x) max", d.amax(dim=(0, 1, 2))[:20].tolist())


def conv_suite(dt):
    cases = [
        dict(cin_segs=[(0, 64)], cout=64, k=3, h=16, w=16, n=1, tag="single tile"),
        dict(cin_segs=[(0, 64)], cout=64, k=1, h=16, w=16, n=1, tag="1x1"),
        dict(cin_segs=[(0, 64)], cout=64, k=3, h=20, w=40, n=2, act="relu"),
        dict(cin_segs=[(0, 64)], cout=64, k=3, h=37, w=52, n=3, act="none", residual=True),
        dict(cin_segs=[(0, 3)], cout=64, k=3, h=24, w=33, n=2, act="lrelu", tag="stem 3ch"),
        dict(cin_segs=[(3, 64), (0, 3)], cout=64, k=3, h=24, w=40, n=2, act="lrelu", tag="cat[lr,feat]"),
        dict(cin_segs=[(0, 64), (64, 64)], cout=64, k=1, h=24, w=40, n=2, act="lrelu", tag="point_conv"),
        dict(cin_segs=[(0, 64)], cout=256, k=3, h=24, w=40, n=1, pixshuf=2, tag="upsample"),
        dict(cin_segs=[(0, 64)], cout=3, k=3, h=24, w=40, n=2, tag="to 3ch"),
        dict(cin_segs=[(0, 8)], cout=32, k=7, h=24, w=40, n=2, act="relu", tag="spynet l0"),
        dict(cin_segs=[(0, 32)], cout=64, k=7, h=24, w=40, n=2, act="relu", tag="spynet l1"),
        dict(cin_segs=[(0, 64)], cout=32, k=7, h=24, w=40, n=2, act="relu", tag="spynet l2"),
        dict(cin_segs=[(0, 32)], cout=16, k=7, h=6, w=10, n=2, act="relu", tag="spynet l3 tiny"),
        dict(cin_segs=[(0, 16)], cout=2, k=7, h=2, w=2, n=3, act="relu", tag="spynet l4 2x2"),
        dict(cin_segs=[(0, 64)], cout=64, k=3, h=24, w=40, n=2, groups=2, act="relu", tag="2 weight groups"),
        dict(cin_segs=[(0, 64)], cout=64, k=3, h=180, w=320, n=4, act="relu", residual=True, tag="cfg3 frame x4"),
    ]
    for i, c in enumerate(cases):
        run(f"conv{i}", lambda c=c, i=i: conv_case(dt, seed=i, **c))


def warp_suite():
    g = torch.Generator().manual_seed(3)
    for dt in (F32, BF16):
        for pad in ("zeros", "border"):
            def f(dt=dt, pad=pad):
                x = torch.randn(2, 64, 19, 27, generator=g)
                fl = (torch.rand(2, 19, 27, 2, generator=g) - 0.5) * 24
                fl[0, 0, 0] = 0
                fl[0, 0, 1] = torch.tensor([1.0, 1.0])
                fl[0, 0, 2] = torch.tensor([0.5, 0.5])
                fl[1, 3, 3] = torch.tensor([500.0, -500.0])
                xr = bf16r(x) if dt == BF16 else x
                ref = O.flow_warp(xr, fl, pad)
                with VF.precision("bf16" if dt == BF16 else "fp32"):
                    got = VF.flow_warp(x.to(dev), fl.to(dev), pad).cpu()
                report(f"flow_warp[{'bf16' if dt == BF16 else 'f32'}] {pad}", (got - ref).abs().max().item(), 3e-2 if dt == BF16 else 2e-5)
            run("warp", f)


def glue_suite():
    def f():
        from conftest import build_state_dict  # type: ignore
    sys.path.insert(0, str(ROOT / "tests"))
    from conftest import build_state_dict
    import numpy as np
    g = np.load(ROOT / "tests" / "golden" / "spynet.npz")
    for tag, kind, tol in (("a", "spynet", 1e-4), ("b", "spynet", 1e-4), ("c", "spynet_amp", 1e-3)):
        def f(tag=tag, kind=kind, tol=tol):
            sp = build_state_dict(kind).to(dev).eval()
            ref, supp = torch.from_numpy(g[f"{tag}_ref"]).to(dev), torch.from_numpy(g[f"{tag}_supp"]).to(dev)
            with torch.no_grad(), VF.precision("fp32"):
                fl = sp(ref, supp).cpu()
            report(f"spynet[f32] golden {tag}", (fl - torch.from_numpy(g[f'{tag}_flow'])).abs().max().item(), tol)
            with torch.no_grad(), VF.precision("bf16"):
                fl = sp(ref, supp).cpu()
            report(f"spynet[bf16] golden {tag}", (fl - torch.from_numpy(g[f'{tag}_flow'])).abs().max().item(), 1e-2 if kind == "spynet" else 0.5)
        run("spynet", f)


def model_suite():
    sys.path.insert(0, str(ROOT / "tests"))
    from conftest import build_state_dict
    import numpy as np
    for name, kind in (("cfg1", "cfg1"), ("ragged", "ragged")):
        def f(name=name, kind=kind):
            g = np.load(ROOT / "tests" / "golden" / f"{name}.npz")
            net = build_state_dict(kind).to(dev).eval()
            sr_ref, lq_ref = torch.from_numpy(g["sr"]), torch.from_numpy(g["lq"])
            for mode, tol in (("fp32", 1e-4), ("bf16", 5e-2)):
                lr = torch.from_numpy(g["lr"]).to(dev)
                with torch.no_grad(), VF.precision(mode):
                    t0 = time.time()
                    sr, lq = net(lr)
                    torch.cuda.synchronize()
                    dtm = time.time() - t0
                report(f"RealBasicVSR[{mode}] {name} sr", (sr.cpu() - sr_ref).abs().max().item(), tol, f"({dtm*1e3:.1f} ms first call)")
                report(f"RealBasicVSR[{mode}] {name} lq", (lq.cpu() - lq_ref).abs().max().item(), tol)
                assert lq.data_ptr() == lr.data_ptr()
                if mode == "bf16":
                    hr = torch.rand(sr_ref.shape, generator=torch.Generator().manual_seed(9))
                    print(f"      PSNR(sr_bf16, sr_ref)={O.psnr(sr.cpu(), sr_ref):.2f} dB; "
                          f"dPSNR vs random HR = {abs(O.psnr(sr.cpu(), hr) - O.psnr(sr_ref, hr)):.5f} dB", flush=True)
        run(name, f)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    print(torch.cuda.get_device_name(0), ops.device_info(), flush=True)
    suites = {
        "warp": warp_suite,
        "conv_f32": lambda: conv_suite(F32),
        "conv_tc": lambda: conv_suite(BF16),
        "glue": glue_suite,
        "model": model_suite,
    }
    for k, fn in suites.items():
        if a.only and k not in a.only.split(","):
            continue
        print(f"=== {k} ===", flush=True)
        try:
            fn()
        except Exception as e:  # noqa: BLE001
            print(f"[EXC] suite {k}: {type(e).__name__}: {e}", flush=True)
            traceback.print_exc()
    bad = [n for n, ok in RESULTS if not ok]
    print(f"SUMMARY: {len(RESULTS) - len(bad)}/{len(RESULTS)} ok; launches={ops.launch_count()}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
