"""First-contact probe of the tcgen05 GEMM on a B200: many shapes/epilogues against an fp64 product."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from counting_detr_b200 import _lib as L

torch.manual_seed(0)
dev = "cuda"
fails = 0


def report(name, got, ref, tol=2e-5):
    global fails
    err = (got.double() - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-30
    rel = err / scale
    ok = rel < tol
    fails += (not ok)
    print(f"{'OK  ' if ok else 'FAIL'} {name}: max_abs_err={err:.3e} rel={rel:.3e}", flush=True)
    return ok


def run_tn(M, N, K, block_n=0, **kw):
    A = torch.randn(M, K, device=dev)
    B = torch.randn(N, K, device=dev)
    ref = A.double() @ B.double().t()
    out = torch.full((M, N), float("nan"), device=dev)
    L.gemm(L.to_split(A), L.to_split(B), M, N, K, mode=0, out_f32=out, block_n=block_n, **kw)
    torch.cuda.synchronize()
    return report(f"TN M={M} N={N} K={K} bn={block_n} {kw}", out, ref)


def run_nt(M, N, K, block_n=0, split_k=1):
    A = torch.randn(K, M, device=dev)
    B = torch.randn(K, N, device=dev)
    ref = A.double().t() @ B.double()
    out = torch.zeros((M, N), device=dev)
    L.gemm(L.to_split(A), L.to_split(B), M, N, K, mode=1, out_f32=out, block_n=block_n, split_k=split_k,
           accumulate=True)
    torch.cuda.synchronize()
    return report(f"NT M={M} N={N} K={K} bn={block_n} split_k={split_k}", out, ref)


print("lib version", L.lib().cdetr_version(), torch.cuda.get_device_name(0), flush=True)
run_tn(128, 128, 64)
run_tn(128, 128, 256)
run_tn(256, 256, 512)
run_tn(128, 16, 64, block_n=16)
run_tn(128, 64, 128, block_n=64)
run_tn(128, 256, 128, block_n=256)
run_tn(300, 200, 200)          # ragged everything (K tail by TMA zero fill)
run_tn(1000, 2, 256)           # head-like N=2
run_tn(4800, 4, 256)
run_tn(16384, 256, 2048)
run_tn(16384, 1024, 256)
run_tn(4096, 2304, 256)
run_nt(128, 128, 64)
run_nt(128, 128, 1024)
run_nt(256, 2304, 4096, split_k=8)
run_nt(256, 256, 16384, split_k=16)
run_nt(200, 72, 1000, split_k=3)  # ragged
run_nt(64, 192, 512)

# precision-policy variants (cdetr_gemm_t.pass_mask, the MASKED kernel instantiations): each must equal the exact product
# of the correspondingly ROUNDED operands (5: b -> bf16, 3: a -> bf16, 1: both), which also proves no term is lost at 7
def run_masked(M, N, K, pm, mode=0, **kw):
    if mode == 0:
        A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev)
    else:
        A = torch.randn(K, M, device=dev); B = torch.randn(K, N, device=dev)
    As, Bs = L.to_split(A), L.to_split(B)
    a_eff = L.from_split(As).double() if pm & 4 else As[0].double()
    b_eff = L.from_split(Bs).double() if pm & 2 else Bs[0].double()
    ref = a_eff @ b_eff.t() if mode == 0 else a_eff.t() @ b_eff
    out = torch.zeros((M, N), device=dev)
    L.gemm(As, Bs, M, N, K, mode=mode, out_f32=out, pass_mask=pm, accumulate=(mode == 1), **kw)
    torch.cuda.synchronize()
    return report(f"masked pm={pm} mode={mode} M={M} N={N} K={K} {kw}", out, ref)


for pm in (5, 3, 1):
    run_masked(384, 256, 320, pm)
    run_masked(16384, 512, 1024, pm)          # CTA-pair shape
    run_masked(256, 2304, 4096, pm, mode=1, split_k=8)

# epilogue: bias + residual(split) + relu + split output; relu-mask; row_scale
M, N, K = 512, 256, 320
A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev)
bias = torch.randn(N, device=dev); res = torch.randn(M, N, device=dev)
ref = torch.relu(A.double() @ B.double().t() + bias.double() + L.from_split(L.to_split(res)).double())
out = torch.empty(M, N, device=dev); outs = torch.zeros(2, M, N, device=dev, dtype=torch.bfloat16)
L.gemm(L.to_split(A), L.to_split(B), M, N, K, out_f32=out, out_split=outs, bias=bias, add_split=L.to_split(res), relu=True)
torch.cuda.synchronize()
report("epilogue bias+res+relu f32", out, ref)
report("epilogue bias+res+relu split", L.from_split(outs), ref)
mask = torch.randn(M, N, device=dev)
rs = torch.randn(M, device=dev)
addf = torch.randn(M, N, device=dev)
ref = ((A.double() @ B.double().t()) * rs.double()[:, None] + addf.double()) * (mask > 0).double()
L.gemm(L.to_split(A), L.to_split(B), M, N, K, out_f32=out, row_scale=rs, add_f32=addf, mask=L.to_split(torch.relu(mask)))
torch.cuda.synchronize()
report("epilogue row_scale+add_f32+mask", out, ref)

# sub-view operands (row slice of a bigger weight, column slice of an activation)
Wbig = torch.randn(1280, 256, device=dev); X = torch.randn(700, 512, device=dev)
Ws = L.to_split(Wbig); Xs = L.to_split(X)
out = torch.empty(700, 256, device=dev)
L.gemm(Xs[:, :, 256:], Ws[:, 512:768], 700, 256, 256, out_f32=out)
torch.cuda.synchronize()
report("subview", out, X[:, 256:].double() @ Wbig[512:768].double().t())

# ---- resident-B schedule (weight slab kept in shared memory; K <= 256, many m-tiles): forced and automatic
os.environ["CDETR_GEMM_RESIDENT"] = "1"
run_tn(4096, 256, 256)
run_tn(5000, 384, 192)         # ragged M / K tail, several slab reloads per CTA
run_tn(40000, 128, 64)
run_tn(3000, 200, 256, block_n=64)
os.environ.pop("CDETR_GEMM_RESIDENT")
run_tn(65536, 512, 128)        # automatic (>= 2 tiles per SM)
M, N, K = 38000, 256, 256
A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev)
bias = torch.randn(N, device=dev); res = torch.randn(M, N, device=dev); mk = torch.randn(M, N, device=dev)
ref = (A.double() @ B.double().t() + bias.double() + L.from_split(L.to_split(res)).double()) * (mk > 0).double()
outs = torch.zeros(2, M, N, device=dev, dtype=torch.bfloat16)
L.gemm(L.to_split(A), L.to_split(B), M, N, K, out_split=outs, bias=bias, add_split=L.to_split(res), mask=L.to_split(torch.relu(mk)))
torch.cuda.synchronize()
report("resident-B auto, bias+res+mask split", L.from_split(outs), ref)

# ---- TMA-staged epilogue on ragged shapes (explicit block_n >= 64 keeps it on the TMA path): column / row tails are
# clipped by the TMA unit, chunks entirely outside the matrix are skipped
run_tn(300, 200, 200, block_n=128)
run_tn(1000, 300, 512, block_n=64)
run_tn(700, 328, 1104, block_n=256)
for (M, N, K, bn) in [(4800, 264, 256, 128), (333, 72, 128, 64), (20000, 512, 64, 256)]:
    A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev)
    bias = torch.randn(N, device=dev); res = torch.randn(M, N, device=dev); mk = torch.randn(M, N, device=dev)
    addf = torch.randn(M, N, device=dev)
    acc = A.double() @ B.double().t()
    outs = torch.zeros(2, M, N, device=dev, dtype=torch.bfloat16); out = torch.full((M, N), float("nan"), device=dev)
    L.gemm(L.to_split(A), L.to_split(B), M, N, K, out_split=outs, out_f32=out, bias=bias, add_split=L.to_split(res),
           mask=L.to_split(torch.relu(mk)), block_n=bn)
    torch.cuda.synchronize()
    ref = (acc + bias.double() + L.from_split(L.to_split(res)).double()) * (mk > 0).double()
    report(f"tma-epi M={M} N={N} K={K} bn={bn} bias+res+mask -> split", L.from_split(outs), ref)
    report(f"tma-epi M={M} N={N} K={K} bn={bn} bias+res+mask -> f32", out, ref)
    out = torch.full((M, N), float("nan"), device=dev)
    L.gemm(L.to_split(A), L.to_split(B), M, N, K, out_f32=out, add_f32=addf, relu=True, block_n=bn)
    torch.cuda.synchronize()
    report(f"tma-epi M={M} N={N} K={K} bn={bn} add_f32+relu -> f32", out, torch.relu(acc + addf.double()))
    out = addf.clone()
    L.gemm(L.to_split(A), L.to_split(B), M, N, K, out_f32=out, accumulate=True, block_n=bn)
    torch.cuda.synchronize()
    report(f"tma-epi M={M} N={N} K={K} bn={bn} accumulate (TMA reduce-add)", out, acc + addf.double())

# ---- implicit 3x3 convolution (TMA shifted windows, no im2col matrix) vs torch conv2d in fp64
import torch.nn.functional as F


def run_conv(Bn, H, W, C, Co, dil):
    x = torch.randn(Bn, C, H, W, device=dev)
    w = torch.randn(Co, C, 3, 3, device=dev) / (3.0 * C ** 0.5)
    scale = torch.rand(Co, device=dev) + 0.5
    dy = torch.randn(Bn, Co, H, W, device=dev)
    xd = x.double().requires_grad_(True); wd_ = (w.double() * scale.double()[:, None, None, None]).requires_grad_(True)
    y_ref = F.conv2d(xd, wd_, padding=dil, dilation=dil)
    y_ref.backward(dy.double())
    Mn = Bn * H * W
    xs = L.to_split(x.permute(0, 2, 3, 1).reshape(Mn, C).contiguous())
    dys = L.to_split(dy.permute(0, 2, 3, 1).reshape(Mn, Co).contiguous())
    wf = torch.zeros(2, Co, 9 * C, device=dev, dtype=torch.bfloat16)
    wdg = torch.zeros(2, C, 9 * Co, device=dev, dtype=torch.bfloat16)
    L.call("cdetr_pack_weight", w.reshape(Co, C, 9).contiguous(), Co, C, 9, scale, wf, None)
    L.call("cdetr_pack_weight_dgrad", w.reshape(Co, C, 9).contiguous(), Co, C, 9, scale, wdg)
    tag = f"conv B={Bn} {H}x{W} C={C}->{Co} dil={dil}"
    y = torch.full((Mn, Co), float("nan"), device=dev)
    L.gemm(xs, wf, Mn, Co, 9 * C, mode=0, out_f32=y, conv=(H, W, C, dil, 1))
    torch.cuda.synchronize()
    report(tag + " fwd", y, y_ref.detach().permute(0, 2, 3, 1).reshape(Mn, Co))
    dx = torch.full((Mn, C), float("nan"), device=dev)
    L.gemm(dys, wdg, Mn, C, 9 * Co, mode=0, out_f32=dx, conv=(H, W, Co, dil, -1))
    torch.cuda.synchronize()
    report(tag + " dgrad", dx, xd.grad.permute(0, 2, 3, 1).reshape(Mn, C))
    dw = torch.zeros(Co, 9 * C, device=dev)
    L.gemm(dys, xs, Co, 9 * C, Mn, mode=1, out_f32=dw, accumulate=True, split_k=4, row_scale=scale,
           block_n=(256 if (9 * C) % 256 == 0 else 128) if C % 128 == 0 else 64, conv=(H, W, C, dil, 1))
    torch.cuda.synchronize()
    # staged layout [Co, (tap, c)]; reference grad is wrt the scaled weight -> multiply by scale for the raw weight
    ref_dw = (wd_.grad * scale.double()[:, None, None, None]).permute(0, 2, 3, 1).reshape(Co, 9 * C)
    # fp32 split-K accumulation over Mn pixels: the noise floor grows with the contraction length (1.3e-5 at 16 k pixels)
    report(tag + " wgrad", dw, ref_dw, 2e-5 if Mn <= 20000 else 6e-5)


run_conv(2, 32, 32, 64, 64, 1)
run_conv(1, 16, 16, 128, 128, 1)
run_conv(2, 32, 32, 256, 256, 2)
run_conv(1, 64, 64, 128, 128, 1)
run_conv(1, 128, 128, 64, 64, 1)
run_conv(3, 32, 32, 512, 512, 2)
# maps whose width does not divide 128 (800 x 800 inputs: 200 / 100 / 50): narrower M tiles (100 of the 128 TMEM lanes),
# generic epilogue; weight gradient with 64-pixel k blocks per image row, zero-filled past W
run_conv(2, 50, 50, 64, 64, 1)
run_conv(1, 50, 50, 256, 256, 2)
run_conv(1, 100, 100, 128, 128, 1)
run_conv(1, 200, 200, 64, 64, 1)
run_conv(2, 12, 20, 64, 128, 1)


def run_conv_fused(Bn, H, W, C, Co, dil):
    """the engine's epilogues on a narrow-tile map: forward split output + ReLU, dgrad split output + ReLU mask"""
    x = torch.randn(Bn, C, H, W, device=dev)
    w = torch.randn(Co, C, 3, 3, device=dev) / (3.0 * C ** 0.5)
    scale = torch.rand(Co, device=dev) + 0.5
    bias = torch.randn(Co, device=dev)
    Mn = Bn * H * W
    xs = L.to_split(x.permute(0, 2, 3, 1).reshape(Mn, C).contiguous())
    wf = torch.zeros(2, Co, 9 * C, device=dev, dtype=torch.bfloat16)
    wdg = torch.zeros(2, C, 9 * Co, device=dev, dtype=torch.bfloat16)
    L.call("cdetr_pack_weight", w.reshape(Co, C, 9).contiguous(), Co, C, 9, scale, wf, None)
    L.call("cdetr_pack_weight_dgrad", w.reshape(Co, C, 9).contiguous(), Co, C, 9, scale, wdg)
    y_ref = torch.relu(F.conv2d(x.double(), w.double() * scale.double()[:, None, None, None], padding=dil, dilation=dil)
                       + bias.double()[None, :, None, None]).permute(0, 2, 3, 1).reshape(Mn, Co)
    ys = torch.zeros(2, Mn, Co, device=dev, dtype=torch.bfloat16)
    L.gemm(xs, wf, Mn, Co, 9 * C, mode=0, out_split=ys, bias=bias, relu=True, conv=(H, W, C, dil, 1))
    torch.cuda.synchronize()
    report(f"conv fused B={Bn} {H}x{W} C={C}->{Co} fwd split+relu", L.from_split(ys), y_ref)
    dy = torch.randn(Mn, Co, device=dev)
    dys = L.to_split(dy)
    mask = L.to_split(torch.randn(Mn, C, device=dev))
    dxs = torch.zeros(2, Mn, C, device=dev, dtype=torch.bfloat16)
    L.gemm(dys, wdg, Mn, C, 9 * Co, mode=0, out_split=dxs, mask=mask, conv=(H, W, Co, dil, -1))
    torch.cuda.synchronize()
    dyn = L.from_split(dys).double().reshape(Bn, H, W, Co).permute(0, 3, 1, 2)
    dx_ref = F.conv_transpose2d(dyn, w.double() * scale.double()[:, None, None, None], padding=dil, dilation=dil)
    dx_ref = dx_ref.permute(0, 2, 3, 1).reshape(Mn, C) * (L.from_split(mask)[:, :C].double() > 0)
    report(f"conv fused B={Bn} {H}x{W} C={C}->{Co} dgrad split+mask", L.from_split(dxs), dx_ref)


run_conv_fused(2, 50, 50, 128, 128, 1)
run_conv_fused(1, 100, 100, 64, 64, 1)

# timing of the epilogue-bound shapes of the C3 step (split output unless noted)
def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for (M, N, K, flav) in [(16384, 1024, 256, "bias+relu"), (16384, 1024, 256, "res+mask"), (16384, 256, 256, "f32"),
                        (4800, 256, 256, "f32"), (262144, 256, 64, "res+relu"), (65536, 512, 128, "res+relu"),
                        (16384, 2048, 512, "res+relu"), (16384, 256, 1024, "f32+addf32"), (65536, 128, 512, "mask")]:
    A = L.to_split(torch.randn(M, K, device=dev)); B = L.to_split(torch.randn(N, K, device=dev))
    outs = torch.empty(2, M, N, device=dev, dtype=torch.bfloat16); outf = torch.empty(M, N, device=dev)
    res = L.to_split(torch.randn(M, N, device=dev)); bias = torch.randn(N, device=dev)
    kw = dict(out_split=outs)
    if flav == "bias+relu": kw.update(bias=bias, relu=True)
    elif flav == "res+mask": kw.update(add_split=res, mask=res)
    elif flav == "res+relu": kw.update(add_split=res, relu=True, bias=bias)
    elif flav == "mask": kw.update(mask=res)
    elif flav == "f32": kw = dict(out_f32=outf, bias=bias)
    elif flav == "f32+addf32": kw = dict(out_f32=outf, add_f32=outf)
    us = timeit(lambda: L.gemm(A, B, M, N, K, **kw))
    print(f"time TN M={M} N={N} K={K} {flav}: {us:.1f} us  {2*M*N*K/us/1e6:.1f} TFLOP/s(fp32-equiv)", flush=True)

# timing (rough): big TN GEMM
for (M, N, K, bn) in [(16384, 2048, 512, 128), (16384, 2048, 512, 256), (16384, 512, 4608, 128), (16384, 512, 4608, 256), (65536, 256, 256, 128), (16384, 1024, 256, 256)]:
    A = L.to_split(torch.randn(M, K, device=dev)); B = L.to_split(torch.randn(N, K, device=dev))
    outs = torch.empty(2, M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        L.gemm(A, B, M, N, K, out_split=outs, block_n=bn)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        L.gemm(A, B, M, N, K, out_split=outs, block_n=bn)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"time TN M={M} N={N} K={K} bn={bn}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s(fp32-equiv) {3*2*M*N*K/ms/1e9:.1f} TFLOP/s(bf16 issued)", flush=True)
print("FAILS", fails)
