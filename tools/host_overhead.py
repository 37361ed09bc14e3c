import time, sys, ctypes as C
sys.path.insert(0, "/root/repo")
import torch
from vsrlab_b200 import ops, _lib as L
from vsrlab_b200._lib import BF16, ACT_RELU
dev = torch.device("cuda:0")
cv = torch.nn.Conv2d(64, 64, 3, 1, 1).to(dev)
pc = ops.PackedConv([cv], [(0, 64)], BF16, 0)
x = torch.randn(1, 8, 8, 64, device=dev).to(torch.bfloat16); y = torch.empty_like(x)
for _ in range(10): ops.conv2d_fwd(pc, [x], [64], 1, 8, 8, act=ACT_RELU, out=y, out_c=64)
torch.cuda.synchronize()
N = 2000
t0 = time.perf_counter()
for _ in range(N): ops.conv2d_fwd(pc, [x], [64], 1, 8, 8, act=ACT_RELU, out=y, out_c=64)
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"ops.conv2d_fwd: {(t1-t0)/N*1e6:.1f} us per call (host)")
# C call only with a prebuilt struct
a = L.ConvArgs(); a.geom = pc.geom; a.inp[0] = x.data_ptr(); a.in_c[0] = 64; a.batch, a.h, a.w = 1, 8, 8
a.imgs_per_group = 1; a.packed = pc.buf.data_ptr(); a.act = ACT_RELU; a.slope = 0.1; a.out = y.data_ptr(); a.out_c = 64
lib = L.load(); st = torch.cuda.current_stream().cuda_stream
t0 = time.perf_counter()
for _ in range(N): lib.vsrb_conv2d_fwd(C.byref(a), st)
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"vsrb_conv2d_fwd C call only: {(t1-t0)/N*1e6:.1f} us per call")
t0 = time.perf_counter()
for _ in range(N): torch.cuda.current_stream().cuda_stream
t1 = time.perf_counter()
print(f"current_stream().cuda_stream: {(t1-t0)/N*1e6:.1f} us")
from vsrlab_b200 import autograd as AG
xx = torch.randn(1, 64, 8, 8, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
for _ in range(5):
    o = AG.conv(cv, [xx], [(0, 64)], "relu"); o.sum().backward()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300):
    o = AG.conv(cv, [xx], [(0, 64)], "relu")
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"autograd conv forward: {(t1-t0)/300*1e6:.1f} us per call")
g = torch.ones_like(o)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300):
    o = AG.conv(cv, [xx], [(0, 64)], "relu"); o.backward(g)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"autograd conv fwd+bwd: {(t1-t0)/300*1e6:.1f} us per call host, {(t2-t0)/300*1e6:.1f} us incl. GPU drain")
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(300):
        o = AG.conv(cv, [xx], [(0, 64)], "relu"); o.backward(g)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"repeat {rep}: fwd+bwd {(t1-t0)/300*1e6:.1f} us host, {(t2-t0)/300*1e6:.1f} us incl. GPU drain")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(300):
    o = AG.conv(cv, [xx], [(0, 64)], "relu"); o.backward(g); xx.grad = None; cv.weight.grad = None; cv.bias.grad = None
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"grads reset each iteration: fwd+bwd {(t1-t0)/300*1e6:.1f} us host, {(t2-t0)/300*1e6:.1f} us incl. GPU drain")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    o = AG.conv(cv, [xx], [(0, 64)], "relu"); o.backward(g)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(8)
# which backward piece is slow on this tiny case?
print("debug_status after autograd loop:", ops.debug_status())
import time as _t
dz = torch.randn(1, 64, 8, 8, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
pt = AG._packed_transposed(cv)
dx = torch.empty_like(dz)
torch.cuda.synchronize(); t0 = _t.perf_counter()
for _ in range(50): ops.conv2d_fwd(pt, [dz], [64], 1, 8, 8, out=dx, out_c=64)
torch.cuda.synchronize(); print(f"dgrad conv: {(_t.perf_counter()-t0)/50*1e6:.1f} us", ops.debug_status())
g = AG._wgrad_geom(cv, ((0, 64),))
dw = torch.zeros_like(cv.weight); db = torch.zeros(64, device=dev)
torch.cuda.synchronize(); t0 = _t.perf_counter()
for _ in range(50): ops.conv2d_wgrad(g, [xx.detach()], [64], dz, 64, 1, 8, 8, 64, dw, db)
torch.cuda.synchronize(); print(f"wgrad: {(_t.perf_counter()-t0)/50*1e6:.1f} us", ops.debug_status())
