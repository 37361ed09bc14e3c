"""Folder-of-PNG video inference pipeline: the B200-side replacement for the I/O loop of the reference's test.py
(test.py:112-141, core/utils.py:282-288).

The reference decodes every frame of a video with PIL in a thread pool into one fp32 tensor, moves fp32 windows to the GPU,
concatenates fp32 outputs on the GPU and encodes PNGs one frame at a time from fp32 tensors (`save_image`: a `make_grid`
copy, mul/add/clamp/permute, a device->host copy and a zlib pass per frame).  At > 1 000 output frames/s per GPU that loop is
the bottleneck by orders of magnitude.  Here:

  * frames are decoded by a thread pool straight into ONE pinned uint8 buffer and cross PCIe as uint8 (4x fewer bytes);
    the `/255` of torchvision's `to_tensor` happens on the device (bit-identical: same fp32 division);
  * windows run through the unchanged public call `model(lr)`; with `out_dtype='uint8'` the last conv's epilogue emits the
    bytes `save_image` would store (floor(clamp(sr,0,1)*255 + 0.5)), so the output crosses PCIe as uint8 as well;
  * the next window's upload and the previous window's download run on a copy stream while the current window computes;
  * PNG encoding runs in a thread pool from the pinned uint8 result (zlib level selectable; level 6 = PIL's default,
    what `save_image` uses).

`upscale_folder` returns per-stage timings so the benchmark can say which stage bounds the end-to-end number."""
from __future__ import annotations

import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import functional as VF


def _decode_into(args):
    path, dst = args
    from PIL import Image
    with Image.open(path) as im:
        a = np.asarray(im.convert("RGB"))
    dst.copy_(torch.from_numpy(np.array(a)).permute(2, 0, 1))              # np.array: a writable copy of PIL's read-only buffer
    return None


def load_video_uint8(folder, pool: ThreadPoolExecutor, pin: bool = True) -> torch.Tensor:
    """[F,3,H,W] uint8 (pinned) of the sorted image files of `folder` (get_video of core/utils.py:285-288, without the
    fp32 blow-up on the host)."""
    from PIL import Image
    paths = sorted(p for p in Path(folder).glob("*") if p.is_file())
    if not paths:
        raise FileNotFoundError(f"no frames under {folder}")
    with Image.open(paths[0]) as im:
        w, h = im.size
    out = torch.empty(len(paths), 3, h, w, dtype=torch.uint8)
    if pin and torch.cuda.is_available():
        out = out.pin_memory()
    list(pool.map(_decode_into, [(p, out[i]) for i, p in enumerate(paths)]))
    return out


def _encode(args):
    frame_u8, path, level = args
    from PIL import Image
    Image.fromarray(frame_u8.permute(1, 2, 0).numpy()).save(path, format="PNG", compress_level=level)


def upscale_folder(model: torch.nn.Module, lr_folder, out_folder: Optional[str] = None, window_size: int = 32, workers: int = 8,
                   device: Optional[torch.device] = None, precision: Optional[str] = None, png_level: int = 6,
                   keep_output: bool = False) -> Dict[str, object]:
    """Super-resolve the PNG frames of `lr_folder` in independent windows of `window_size` frames (test.py:125-131) and,
    if `out_folder` is given, write `img%05d.png` files there (test.py:138-141).  Returns timings and, with
    `keep_output`, the uint8 result [F,3,sH,sW] (pinned host tensor)."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    pool = ThreadPoolExecutor(workers)
    t0 = time.perf_counter()
    video = load_video_uint8(lr_folder, pool)
    t_decode = time.perf_counter() - t0
    F_, _, h, w = video.shape
    windows = [(i, min(i + window_size, F_)) for i in range(0, F_, window_size)]
    copy_stream = torch.cuda.Stream(device=device)
    main = torch.cuda.current_stream(device)
    out_host: Optional[torch.Tensor] = None
    pending: List = []
    futures = []
    if out_folder:
        Path(out_folder).mkdir(parents=True, exist_ok=True)
    t1 = time.perf_counter()

    def upload(a, b):
        with torch.cuda.stream(copy_stream):
            t = video[a:b].to(device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return t, ev

    # a 0-dim DEVICE divisor: torch multiplies by the reciprocal when dividing by a host scalar, which is not the correctly
    # rounded division `to_tensor` performs on the CPU
    d255 = torch.tensor(255.0, device=device)
    nxt = upload(*windows[0])
    ctx_p = VF.precision(precision) if precision else None
    if ctx_p:
        ctx_p.__enter__()
    try:
        with torch.no_grad(), VF.output_dtype("uint8"):
            for k, (a, b) in enumerate(windows):
                u8, ready = nxt
                main.wait_event(ready)
                u8.record_stream(main)
                lr = torch.div(u8.float(), d255).unsqueeze(0)          # torchvision to_tensor: uint8 -> fp32 / 255, bit for bit
                if k + 1 < len(windows):
                    nxt = upload(*windows[k + 1])
                sr, _ = model(lr)                                      # [1, T, 3, sH, sW] uint8
                if out_host is None:
                    out_host = torch.empty(F_, 3, sr.shape[-2], sr.shape[-1], dtype=torch.uint8).pin_memory()
                done = torch.cuda.Event()
                done.record(main)
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(done)
                    sr.record_stream(copy_stream)
                    out_host[a:b].copy_(sr[0], non_blocking=True)
                    landed = torch.cuda.Event()
                    landed.record(copy_stream)
                pending.append((a, b, landed))
                # encode windows whose download has finished while the GPU keeps going
                while pending and (pending[0][2].query() or k + 1 == len(windows)):
                    pa, pb, ev = pending.pop(0)
                    ev.synchronize()
                    if out_folder:
                        futures += [pool.submit(_encode, (out_host[i], str(Path(out_folder) / f"img{i:05d}.png"), png_level))
                                    for i in range(pa, pb)]
    finally:
        if ctx_p:
            ctx_p.__exit__(None, None, None)
    torch.cuda.synchronize(device)
    t_gpu = time.perf_counter() - t1
    for f in futures:
        f.result()
    t_total = time.perf_counter() - t0
    pool.shutdown()
    res: Dict[str, object] = {"frames": F_, "lr_size": (h, w), "windows": len(windows), "decode_s": t_decode, "infer_and_copy_s": t_gpu,
                              "total_s": t_total, "frames_per_s": F_ / t_total}
    if keep_output:
        res["output"] = out_host
    return res
