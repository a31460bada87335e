"""GPU probe of every non-GEMM kernel against torch (fp32, TF32 off) / the oracle.  Run via gpurun."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
from counting_detr_b200 import _lib as L
from oracle import criterion as OC, model as OM

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = "cuda"
torch.manual_seed(0)
fails = []


def report(name, got, ref, tol=2e-5):
    got = got.double(); ref = ref.double()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-30
    ok = err / scale < tol and not torch.isnan(got).any().item()
    if not ok:
        fails.append(name)
    print(f"{'OK  ' if ok else 'FAIL'} {name}: max_abs_err={err:.3e} rel={err/scale:.3e}", flush=True)


def S(x):  # NHWC/rows fp32 -> split [2, rows, C]
    return L.to_split(x.reshape(-1, x.shape[-1]).contiguous())


def zs(rows, cols):
    return torch.zeros(2, rows, cols, device=dev, dtype=torch.bfloat16)


# ---------------------------------------------------------------- backbone layout kernels
B, H, W, C = 2, 12, 10, 16
x = torch.randn(B, H, W, C, device=dev)
xs = S(x)
xn = L.from_split(xs).reshape(B, H, W, C)          # the values the kernels actually see
for stride, dil in [(1, 1), (2, 1), (1, 2)]:
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    col = zs(B * Ho * Wo, 9 * C)
    L.call("cdetr_im2col3x3", xs, B, H, W, C, stride, dil, col)
    ref = F.unfold(xn.permute(0, 3, 1, 2), 3, dilation=dil, padding=dil, stride=stride)  # [B, C*9, L]
    ref = ref.view(B, C, 9, Ho * Wo).permute(0, 3, 2, 1).reshape(B * Ho * Wo, 9 * C)
    report(f"im2col s{stride} d{dil}", L.from_split(col), ref, 1e-7)
    dcol = torch.randn(B * Ho * Wo, 9 * C, device=dev)
    dcs = L.to_split(dcol); dcn = L.from_split(dcs)
    dx = zs(B * H * W, C)
    L.call("cdetr_col2im3x3", dcs, B, H, W, C, stride, dil, xs_mask := S(torch.relu(x)), dx)
    refdx = F.fold(dcn.view(B, Ho * Wo, 9, C).permute(0, 3, 2, 1).reshape(B, C * 9, Ho * Wo), (H, W), 3,
                   dilation=dil, padding=dil, stride=stride).permute(0, 2, 3, 1)
    refdx = refdx * (L.from_split(xs_mask).reshape(B, H, W, C) > 0)
    report(f"col2im s{stride} d{dil}", L.from_split(dx).reshape(B, H, W, C), refdx, 1e-5)
# fused stem (conv 7x7 s2 p3 + FrozenBN + ReLU, no im2col matrix) vs torch conv2d in fp64, odd sizes / partial patches
for (Bs, Hs, Ws) in [(2, 64, 96), (1, 70, 90), (3, 33, 47), (1, 256, 256)]:
    img = torch.randn(Bs, 3, Hs, Ws, device=dev)
    w7 = torch.randn(64, 3, 7, 7, device=dev) * 0.1
    sc = torch.rand(64, device=dev) + 0.5
    sh = torch.randn(64, device=dev)
    wp = torch.zeros(64, 152, device=dev)
    wp[:, :147] = (w7 * sc[:, None, None, None]).permute(0, 2, 3, 1).reshape(64, 147)      # k = (r*7 + s)*3 + c
    wps = L.to_split(wp)
    Hs0, Ws0 = (Hs + 6 - 7) // 2 + 1, (Ws + 6 - 7) // 2 + 1
    o = zs(Bs * Hs0 * Ws0, 64)
    L.call("cdetr_stem_conv", img, Bs, Hs, Ws, wps, sh, o)
    wn = L.from_split(wps)[:, :147].reshape(64, 7, 7, 3).permute(0, 3, 1, 2)
    ref = torch.relu(F.conv2d(img.double(), wn.double(), stride=2, padding=3) + sh.double()[None, :, None, None])
    report(f"stem_conv B{Bs} {Hs}x{Ws}", L.from_split(o).reshape(Bs, Hs0, Ws0, 64), ref.permute(0, 2, 3, 1), 2e-5)
y = zs(B * 6 * 5, C)
L.call("cdetr_maxpool3x3s2", xs, B, H, W, C, y)
report("maxpool", L.from_split(y).reshape(B, 6, 5, C), F.max_pool2d(xn.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1), 1e-7)
y = zs(B * 6 * 5, C)
L.call("cdetr_subsample2", xs, B, H, W, C, y)
report("subsample2", L.from_split(y).reshape(B, 6, 5, C), xn[:, ::2, ::2], 1e-7)
up = zs(B * H * W, C)
L.call("cdetr_upsample2_zero", y, B, H, W, C, up)
refu = torch.zeros_like(xn); refu[:, ::2, ::2] = xn[:, ::2, ::2]
report("upsample2_zero", L.from_split(up).reshape(B, H, W, C), refu, 1e-7)
img = torch.randn(2, 3, 20, 24, device=dev)
col = zs(2 * 10 * 12, 152)
L.call("cdetr_stem_im2col", img, 2, 20, 24, col)
ref = F.unfold(img, 7, padding=3, stride=2).view(2, 3, 49, 120).permute(0, 3, 2, 1).reshape(240, 147)
report("stem_im2col", L.from_split(col)[:, :147], ref, 2e-5)
report("stem_im2col pad", L.from_split(col)[:, 147:], torch.zeros(240, 5, device=dev) + 0, 1)
# weight pack / conv through GEMM
w = torch.randn(24, C, 3, 3, device=dev)
scale = torch.rand(24, device=dev) + 0.5
wp, wpt = zs(24, 9 * C), zs(9 * C, 24)
L.call("cdetr_pack_weight", w, 24, C, 9, scale, wp, wpt)
refw = (w * scale[:, None, None, None]).permute(0, 2, 3, 1).reshape(24, 9 * C)
report("pack_weight", L.from_split(wp), refw, 2e-5)
report("pack_weight_T", L.from_split(wpt), refw.t(), 2e-5)
col = zs(B * H * W, 9 * C)
L.call("cdetr_im2col3x3", xs, B, H, W, C, 1, 1, col)
out = torch.empty(B * H * W, 24, device=dev)
L.gemm(col, wp, B * H * W, 24, 9 * C, out_f32=out)
refc = F.conv2d(xn.permute(0, 3, 1, 2).double(), (w * scale[:, None, None, None]).double(), padding=1).permute(0, 2, 3, 1)
report("conv3x3 via im2col+gemm", out.view(B, H, W, 24), refc, 2e-5)
g = torch.randn(24, 9 * C, device=dev); grad = torch.randn(24, C, 3, 3, device=dev); g0 = grad.clone()
L.call("cdetr_unpack_conv_grad", g, 24, C, 9, grad)
report("unpack_conv_grad", grad, g0 + g.view(24, 9, C).permute(0, 2, 1).reshape(24, C, 3, 3), 1e-6)
bw, bb, rm, rv = [torch.rand(24, device=dev) + 0.5 for _ in range(4)]
sc, sh = torch.empty(24, device=dev), torch.empty(24, device=dev)
L.call("cdetr_bn_fold", bw, bb, rm, rv, 1e-5, 24, sc, sh)
report("bn_fold scale", sc, bw * (rv + 1e-5).rsqrt(), 1e-6)
report("bn_fold shift", sh, bb - rm * bw * (rv + 1e-5).rsqrt(), 1e-6)
# exemplar concat fwd/bwd vs autograd
Cx = 32
xe = torch.randn(B, H, W, Cx, device=dev); xes = S(xe); xen = L.from_split(xes).reshape(B, H, W, Cx).requires_grad_()
yx = torch.tensor([[3, 4], [7, 2], [3, 4]], dtype=torch.int32, device=dev)
p_out = torch.empty(B, Cx, device=dev); cat = zs(B * H * W, 2 * Cx)
L.call("cdetr_exemplar_concat", xes, B, H, W, Cx, yx, 3, p_out, cat)
pr = torch.stack([xen[:, int(a), int(b_)] for a, b_ in yx.tolist()]).mean(0)
refcat = torch.cat([xen, xen * pr[:, None, None, :]], -1)
report("exemplar_concat", L.from_split(cat).reshape(B, H, W, 2 * Cx), refcat.detach(), 2e-5)
dcat = torch.randn(B, H, W, 2 * Cx, device=dev); dcs = S(dcat); dcn = L.from_split(dcs).reshape(B, H, W, 2 * Cx)
refcat.backward(dcn)
dxe = zs(B * H * W, Cx); dp = torch.empty(B, Cx, device=dev)
L.call("cdetr_exemplar_concat_bwd", dcs, xes, p_out, B, H, W, Cx, yx, 3, dp, None, dxe)
report("exemplar_concat_bwd", L.from_split(dxe).reshape(B, H, W, Cx), xen.grad, 3e-5)

# ---------------------------------------------------------------- norms
Bn, N, E = 3, 40, 256
x = torch.randn(Bn, N, E, device=dev, requires_grad=True)
gam = (torch.rand(E, device=dev) + 0.5).requires_grad_(); bet = torch.randn(E, device=dev).requires_grad_()
y = torch.empty(Bn, N, E, device=dev); ysp = zs(Bn * N, E); st = torch.empty(Bn * 32 * 2, device=dev)
L.call("cdetr_groupnorm_fwd", x, Bn, N, E, 32, gam, bet, 1e-5, y, ysp, st)
ref = F.group_norm(x.permute(0, 2, 1), 32, gam, bet, 1e-5).permute(0, 2, 1)
report("groupnorm_fwd", y, ref.detach(), 1e-5)
report("groupnorm_fwd split", L.from_split(ysp).view(Bn, N, E), ref.detach(), 1e-5)
dy = torch.randn(Bn, N, E, device=dev)
ref.backward(dy)
dx = torch.empty(Bn, N, E, device=dev); dg = torch.zeros(E, device=dev); db = torch.zeros(E, device=dev)
L.call("cdetr_groupnorm_bwd", dy, x, Bn, N, E, 32, gam, st, dx, None, dg, db)
report("groupnorm_bwd dx", dx, x.grad, 2e-5); report("groupnorm_bwd dgamma", dg, gam.grad, 2e-5); report("groupnorm_bwd dbeta", db, bet.grad, 2e-5)
M = 333
x = torch.randn(M, E, device=dev, requires_grad=True); r = torch.randn(M, E, device=dev, requires_grad=True)
gam = (torch.rand(E, device=dev) + 0.5).requires_grad_(); bet = torch.randn(E, device=dev).requires_grad_()
z = torch.empty(M, E, device=dev); y = torch.empty(M, E, device=dev); ysp = zs(M, E); st = torch.empty(M * 2, device=dev)
L.call("cdetr_layernorm_fwd", x, r, M, E, gam, bet, 1e-5, z, y, ysp, st)
ref = F.layer_norm(x + r, (E,), gam, bet, 1e-5)
report("layernorm_fwd", y, ref.detach(), 1e-5); report("layernorm_fwd split", L.from_split(ysp), ref.detach(), 1e-5)
dy = torch.randn(M, E, device=dev); dy2 = torch.randn(M, E, device=dev)
ref.backward(dy + dy2)
dz = torch.empty(M, E, device=dev); dzs = zs(M, E); dg = torch.zeros(E, device=dev); db = torch.zeros(E, device=dev)
L.call("cdetr_layernorm_bwd", dy, dy2, z, st, M, E, gam, dz, dzs, dg, db)
report("layernorm_bwd dz", dz, x.grad, 2e-5); report("layernorm_bwd dz split", L.from_split(dzs), x.grad, 2e-5)
report("layernorm_bwd dgamma", dg, gam.grad, 2e-5); report("layernorm_bwd dbeta", db, bet.grad, 2e-5)

# ---------------------------------------------------------------- transformer glue
pos = torch.rand(77, device=dev).requires_grad_()
out = torch.empty(77, 256, device=dev)
L.call("cdetr_sine_embed", pos, 77, 1, 256, 0, 256, out)
ref = OM.sine_embed_1d(pos.detach().cpu()).to(dev)
report("sine_embed_1d", out, ref, 2e-5)
p2 = torch.rand(50, 2, device=dev)
out2 = torch.empty(50, 256, device=dev)
L.call("cdetr_sine_embed", p2[:, 1:], 50, 2, 128, 0, 256, out2)     # y first
L.call("cdetr_sine_embed", p2, 50, 2, 128, 128, 256, out2)          # then x
report("sine_embed_2d", out2, OM.sine_embed_2d(p2.cpu()).to(dev), 2e-5)
pd = pos.detach().clone().double().requires_grad_()
i = torch.arange(256, device=dev, dtype=torch.float64); dt = 10000.0 ** (2 * (i // 2) / 256)
v = pd[:, None] * (2 * np.pi) / dt
emb = torch.stack((v[:, 0::2].sin(), v[:, 1::2].cos()), -1).flatten(-2)
de = torch.randn(77, 256, device=dev)
emb.backward(de.double())
dpos = torch.zeros(77, device=dev)
L.call("cdetr_sine_embed_bwd", pos, 77, 1, 256, 0, 256, de, dpos)
report("sine_embed_bwd", dpos, pd.grad, 1e-4)
Bq, Hq, Wq = 2, 5, 7
src = torch.randn(Bq, Hq, Wq, E, device=dev); per = torch.randn(Bq, Wq, E, device=dev); pec = torch.randn(Bq, Hq, E, device=dev)
o1 = torch.empty(Bq * Hq * Wq, E, device=dev); o1s = zs(Bq * Hq * Wq, E)
L.call("cdetr_add_bcast", src, per, Bq * Hq * Wq, E, 1, Hq, Wq, 0, o1, o1s)
report("add_bcast row", o1.view(Bq, Hq, Wq, E), src + per[:, None], 1e-6); report("add_bcast row split", L.from_split(o1s).view(Bq, Hq, Wq, E), src + per[:, None], 1e-5)
L.call("cdetr_add_bcast", src, pec, Bq * Hq * Wq, E, 2, Hq, Wq, 0, o1, None)
report("add_bcast col", o1.view(Bq, Hq, Wq, E), src + pec[:, :, None], 1e-6)
L.call("cdetr_add_bcast", src, None, Bq * Hq * Wq, E, 0, Hq, Wq, 0, None, o1s)
report("add_bcast none->split", L.from_split(o1s).view(Bq, Hq, Wq, E), src, 1e-5)
r1 = torch.empty(Bq, Wq, E, device=dev); r1s = zs(Bq * Wq, E)
L.call("cdetr_reduce_axis", src, Bq, Hq, Wq, E, 1, 1.0 / Hq, per, 0, r1, r1s)
report("reduce_axis h", r1, src.mean(1) + per, 1e-6); report("reduce_axis h split", L.from_split(r1s).view(Bq, Wq, E), src.mean(1) + per, 1e-5)
r2 = torch.ones(Bq, Hq, E, device=dev)
L.call("cdetr_reduce_axis", src, Bq, Hq, Wq, E, 2, 1.0, None, 1, r2, None)
report("reduce_axis w acc", r2, src.sum(2) + 1, 1e-6)
a_, b_, c_ = [torch.randn(Bq * Hq * Wq, E, device=dev) for _ in range(3)]
oc = torch.empty(Bq * Hq * Wq, E, device=dev)
L.call("cdetr_combine_bcast", a_, b_, c_, per, 0.5, pec, 0.25, Bq * Hq * Wq, E, Hq, Wq, oc)
report("combine_bcast", oc.view(Bq, Hq, Wq, E), (a_ + b_ + c_).view(Bq, Hq, Wq, E) + 0.5 * per[:, None] + 0.25 * pec[:, :, None], 1e-6)
xx = torch.randn(1000, 300, device=dev); cs = torch.ones(300, device=dev)
L.call("cdetr_colsum", xx, None, 300, 1000, 300, cs)
report("colsum f32", cs, xx.sum(0) + 1, 1e-5)
xxs = L.to_split(xx); cs = torch.zeros(300, device=dev)
L.call("cdetr_colsum", None, xxs, 0, 1000, 300, cs)
report("colsum split", cs, L.from_split(xxs).sum(0), 1e-5)
t = torch.randn(90, 4, device=dev, requires_grad=True); ref_pts = torch.rand(90, 2, device=dev).requires_grad_()
bx = torch.empty(90, 4, device=dev)
L.call("cdetr_box_head_fwd", t, ref_pts, 90, bx)
inv = OM.inverse_sigmoid(ref_pts)
refb = torch.cat([t[:, :2] + inv, t[:, 2:]], -1).sigmoid()
report("box_head_fwd", bx, refb.detach(), 1e-5)
db_ = torch.randn(90, 4, device=dev); refb.backward(db_)
dt_ = torch.empty(90, 4, device=dev); dts = zs(90, 8); dref = torch.zeros(90, 2, device=dev)
L.call("cdetr_box_head_bwd", db_, bx, ref_pts, 90, dt_, dts, dref)
report("box_head_bwd dt", dt_, t.grad, 1e-5); report("box_head_bwd dt split", L.from_split(dts)[:, :4], t.grad, 2e-5); report("box_head_bwd dref", dref, ref_pts.grad, 1e-4)

# ---------------------------------------------------------------- attention cores
def rcda_ref(qr, qc, kr, kc, v, nh):
    Bz, Lq, E_ = qr.shape; Hh, Ww = v.shape[1:3]; d = E_ // nh
    s_r = torch.einsum("blnd,bwnd->bnlw", qr.view(Bz, Lq, nh, d) * d ** -0.5, kr.view(Bz, Ww, nh, d))
    s_c = torch.einsum("blnd,bhnd->bnlh", qc.view(Bz, Lq, nh, d) * d ** -0.5, kc.view(Bz, Hh, nh, d))
    a_r, a_c = s_r.softmax(-1), s_c.softmax(-1)
    tt = torch.einsum("bnlh,bhwnd->bnlwd", a_c, v.view(Bz, Hh, Ww, nh, d))
    return torch.einsum("bnlw,bnlwd->blnd", a_r, tt).reshape(Bz, Lq, E_), a_r, a_c

for (Bz, Lq, Hh, Ww) in [(2, 300, 32, 32), (1, 70, 9, 13), (1, 1024, 32, 32), (1, 130, 50, 50), (2, 500, 50, 50),
                         (1, 2500, 50, 50), (1, 300, 38, 64), (1, 257, 64, 33), (1, 90, 20, 40)]:
    nh = 8
    qr, qc = [(torch.randn(Bz, Lq, E, device=dev) * 1.5).requires_grad_() for _ in range(2)]
    kr = (torch.randn(Bz, Ww, E, device=dev)).requires_grad_(); kc = torch.randn(Bz, Hh, E, device=dev).requires_grad_()
    v = torch.randn(Bz, Hh, Ww, E, device=dev).requires_grad_()
    ar = torch.empty(Bz, nh, Ww, Lq, device=dev); ac = torch.empty(Bz, nh, Hh, Lq, device=dev); o = zs(Bz * Lq, E)
    L.call("cdetr_rcda_fwd", Bz, Lq, Hh, Ww, E, nh, qr, qc, kr, kc, v, None, None, ar, ac, o)
    ref, a_r, a_c = rcda_ref(qr, qc, kr, kc, v, nh)
    tag = f"B{Bz} L{Lq} {Hh}x{Ww}"
    report(f"rcda_fwd {tag}", L.from_split(o).view(Bz, Lq, E), ref.detach(), 2e-5)
    report(f"rcda_fwd A_r {tag}", ar.permute(0, 1, 3, 2), a_r.detach(), 1e-5)
    if Hh <= 64 and Ww <= 64:
        ar2 = torch.empty_like(ar); ac2 = torch.empty_like(ac); o2 = zs(Bz * Lq, E)
        L.call("cdetr_rcda_fwd_tc", Bz, Lq, Hh, Ww, E, nh, qr, qc, kr, kc, S(v.detach()), None, None, ar2, ac2, o2)
        report(f"rcda_fwd_tc {tag}", L.from_split(o2).view(Bz, Lq, E), ref.detach(), 3e-5)
        # logits come from the 3-pass split-bf16 tensor-core product (2^-16 relative per operand pair): 3e-5, like O
        report(f"rcda_fwd_tc A_r {tag}", ar2.permute(0, 1, 3, 2), a_r.detach(), 3e-5)
        report(f"rcda_fwd_tc A_c {tag}", ac2.permute(0, 1, 3, 2), a_c.detach(), 3e-5)
    dO = torch.randn(Bz, Lq, E, device=dev)
    ref.backward(dO)
    dsr = torch.empty_like(ar); dsc = torch.empty_like(ac)
    dqr, dqc, dkr, dkc, dv = zs(Bz * Lq, E), zs(Bz * Lq, E), zs(Bz * Ww, E), zs(Bz * Hh, E), zs(Bz * Hh * Ww, E)
    L.call("cdetr_rcda_bwd", Bz, Lq, Hh, Ww, E, nh, qr, qc, kr, kc, v, ar, ac, dO, dsr, dsc, dqr, dqc, dkr, dkc, dv)
    if Hh <= 64 and Ww <= 64:
        dsr2 = torch.empty_like(ar); dsc2 = torch.empty_like(ac)
        dqr2, dqc2, dkr2, dkc2, dv2 = zs(Bz * Lq, E), zs(Bz * Lq, E), zs(Bz * Ww, E), zs(Bz * Hh, E), zs(Bz * Hh * Ww, E)
        L.call("cdetr_rcda_bwd_q_tc", Bz, Lq, Hh, Ww, E, nh, kr, kc, S(v.detach()), ar, ac, dO, dsr2, dsc2, dqr2, dqc2)
        L.call("cdetr_rcda_bwd_kv", Bz, Lq, Hh, Ww, E, nh, qr, qc, ar, ac, dO, dsr2, dsc2, dkr2, dkc2, dv2)
        dv3 = zs(Bz * Hh * Ww, E); dkr3, dkc3 = zs(Bz * Ww, E), zs(Bz * Hh, E)
        L.call("cdetr_rcda_bwd_v_tc", Bz, Lq, Hh, Ww, E, nh, ar, ac, S(dO), dv3)
        L.call("cdetr_rcda_bwd_k", Bz, Lq, Hh, Ww, E, nh, qr, qc, dsr2, dsc2, dkr3, dkc3)
        report(f"rcda_bwd_v_tc dv {tag}", L.from_split(dv3).view_as(v), v.grad, 5e-5)
        report(f"rcda_bwd_k dkr {tag}", L.from_split(dkr3).view_as(kr), kr.grad, 5e-5)
        report(f"rcda_bwd_k dkc {tag}", L.from_split(dkc3).view_as(kc), kc.grad, 5e-5)
        report(f"rcda_bwd_tc dqr {tag}", L.from_split(dqr2).view_as(qr), qr.grad, 5e-5)
        report(f"rcda_bwd_tc dqc {tag}", L.from_split(dqc2).view_as(qc), qc.grad, 5e-5)
        report(f"rcda_bwd_tc dsr {tag}", dsr2, dsr, 5e-5)
        report(f"rcda_bwd_tc dsc {tag}", dsc2, dsc, 5e-5)
        report(f"rcda_bwd_tc dkr {tag}", L.from_split(dkr2).view_as(kr), kr.grad, 5e-5)
        report(f"rcda_bwd_tc dkc {tag}", L.from_split(dkc2).view_as(kc), kc.grad, 5e-5)
        report(f"rcda_bwd_tc dv {tag}", L.from_split(dv2).view_as(v), v.grad, 5e-5)
    report(f"rcda_bwd dqr {tag}", L.from_split(dqr).view_as(qr), qr.grad, 5e-5)
    report(f"rcda_bwd dqc {tag}", L.from_split(dqc).view_as(qc), qc.grad, 5e-5)
    report(f"rcda_bwd dkr {tag}", L.from_split(dkr).view_as(kr), kr.grad, 5e-5)
    report(f"rcda_bwd dkc {tag}", L.from_split(dkc).view_as(kc), kc.grad, 5e-5)
    report(f"rcda_bwd dv {tag}", L.from_split(dv).view_as(v), v.grad, 5e-5)
# masked RCDA (resident-V kernel: 6x7; streaming kernel: 40x50)
for (Hh, Ww, hv, wv) in [(6, 7, 4, 5), (40, 50, 33, 41)]:
    Bz, Lq, nh = 1, 140, 8
    qr, qc = [torch.randn(Bz, Lq, E, device=dev) for _ in range(2)]
    kr, kc, v = torch.randn(Bz, Ww, E, device=dev), torch.randn(Bz, Hh, E, device=dev), torch.randn(Bz, Hh, Ww, E, device=dev)
    mr = torch.zeros(Bz, Ww, dtype=torch.uint8, device=dev); mr[:, wv:] = 1
    mc = torch.zeros(Bz, Hh, dtype=torch.uint8, device=dev); mc[:, hv:] = 1
    ar = torch.empty(Bz, nh, Ww, Lq, device=dev); ac = torch.empty(Bz, nh, Hh, Lq, device=dev); o = zs(Bz * Lq, E)
    L.call("cdetr_rcda_fwd", Bz, Lq, Hh, Ww, E, nh, qr, qc, kr, kc, v, mr, mc, ar, ac, o)
    ref, _, _ = rcda_ref(qr, qc, kr[:, :wv], kc[:, :hv], v[:, :hv, :wv], nh)
    report(f"rcda_fwd masked {Hh}x{Ww}", L.from_split(o).view(Bz, Lq, E), ref, 2e-5)
    o2 = zs(Bz * Lq, E)
    L.call("cdetr_rcda_fwd_tc", Bz, Lq, Hh, Ww, E, nh, qr, qc, kr, kc, S(v), mr, mc, ar, ac, o2)
    report(f"rcda_fwd_tc masked {Hh}x{Ww}", L.from_split(o2).view(Bz, Lq, E), ref, 3e-5)

for (Bz, Lq) in [(2, 300), (1, 77), (1, 500)]:
    nh = 8; d = 32
    qkv = torch.randn(Bz, Lq, 3 * E, device=dev, requires_grad=True)
    q, k, v = qkv[..., :E], qkv[..., E:2 * E], qkv[..., 2 * E:]
    o = zs(Bz * Lq, E); lse = torch.empty(Bz, nh, Lq, device=dev)
    L.call("cdetr_mha_fwd", Bz, Lq, E, nh, q, k, v, 3 * E, o, lse)
    s = torch.einsum("bind,bjnd->bnij", q.reshape(Bz, Lq, nh, d) * d ** -0.5, k.reshape(Bz, Lq, nh, d))
    ref = torch.einsum("bnij,bjnd->bind", s.softmax(-1), v.reshape(Bz, Lq, nh, d)).reshape(Bz, Lq, E)
    report(f"mha_fwd B{Bz} L{Lq}", L.from_split(o).view(Bz, Lq, E), ref.detach(), 2e-5)
    dO = torch.randn(Bz, Lq, E, device=dev); ref.backward(dO)
    dsum = torch.empty(Bz, nh, Lq, device=dev); dq, dk, dv = zs(Bz * Lq, E), zs(Bz * Lq, E), zs(Bz * Lq, E)
    L.call("cdetr_mha_bwd", Bz, Lq, E, nh, q, k, v, 3 * E, o, lse, dO, dsum, dq, dk, dv)
    g = qkv.grad
    report(f"mha_bwd dq L{Lq}", L.from_split(dq).view(Bz, Lq, E), g[..., :E], 5e-5)
    report(f"mha_bwd dk L{Lq}", L.from_split(dk).view(Bz, Lq, E), g[..., E:2 * E], 5e-5)
    report(f"mha_bwd dv L{Lq}", L.from_split(dv).view(Bz, Lq, E), g[..., 2 * E:], 5e-5)

# ---------------------------------------------------------------- matcher + losses
from scipy.optimize import linear_sum_assignment as lsa
def run_match(Bm, Q, Ts, dup=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(Bm, Q, 2, generator=g)
    boxes = torch.cat([torch.rand(Bm, Q, 2, generator=g), torch.rand(Bm, Q, 2, generator=g) * 0.2 + 0.01], -1)
    tg = [torch.cat([torch.rand(T, 2, generator=g), torch.rand(T, 2, generator=g) * 0.2 + 0.01], -1) for T in Ts]
    if dup:
        tg = [torch.cat([t[: max(1, len(t) // 2)]] * 2)[: len(t)] if len(t) > 1 else t for t in tg]
    Tmax = max(Ts)
    off = torch.tensor(np.concatenate([[0], np.cumsum(Ts)]), dtype=torch.int32, device=dev)
    tcat = torch.cat(tg).to(dev) if sum(Ts) else torch.zeros(1, 4, device=dev)
    cost = torch.zeros(Bm, Q * max(Tmax, 1), device=dev)
    K = min(Q, Tmax)
    oq = torch.full((Bm, max(K, 1)), -1, dtype=torch.int64, device=dev); ot = oq.clone()
    on = torch.zeros(Bm, dtype=torch.int32, device=dev); status = torch.zeros(1, dtype=torch.int32, device=dev)
    lg, bx = logits.to(dev), boxes.to(dev)
    L.call("cdetr_match_cost", lg, 2, bx, tcat, off, Bm, Q, Tmax, 2.0, 5.0, 2.0, cost)
    L.call("cdetr_lsap", cost, off, Bm, Q, Tmax, oq, ot, on, status)
    torch.cuda.synchronize()
    bad = 0; maxc = 0.0
    for b in range(Bm):
        T = Ts[b]
        if T == 0:
            bad += int(on[b].item() != 0); continue
        cref = OC.match_cost(logits[b], boxes[b], tg[b])
        cg = cost[b, : Q * T].cpu()
        cg = cg.view(T, Q).t() if T < Q else cg.view(Q, T)
        maxc = max(maxc, (cg - cref).abs().max().item())
        # solver exactness on the SAME (device-produced) costs, and end-to-end vs reference costs
        i1, j1 = lsa(cg.numpy()); i2, j2 = lsa(cref.numpy())
        n = on[b].item()
        gq, gt = oq[b, :n].cpu().numpy(), ot[b, :n].cpu().numpy()
        same_cost = np.array_equal(gq, i1) and np.array_equal(gt, j1)
        e2e = np.array_equal(gq, i2) and np.array_equal(gt, j2)
        bad += (not same_cost) + (not e2e)
    ok = bad == 0 and status.item() == 0
    if not ok: fails.append(f"match {Bm}x{Q}x{Ts[:3]}")
    print(f"{'OK  ' if ok else 'FAIL'} matcher B={Bm} Q={Q} T={Ts[:4]} dup={dup}: mismatches={bad} max|cost-ref|={maxc:.2e}", flush=True)
    return lg, bx, tg, off, tcat, oq, ot, on

run_match(4, 300, [50, 50, 50, 50])
run_match(5, 300, [50, 1, 0, 17, 300], seed=1)
run_match(3, 50, [80, 50, 7], seed=2)
run_match(4, 300, [50, 50, 50, 50], dup=True, seed=3)
run_match(2, 500, [100, 100], seed=4)
run_match(2, 1000, [1000, 1000], seed=5)
run_match(1, 1000, [1000], dup=True, seed=6)
run_match(2, 600, [700, 50], seed=7)

# losses vs oracle (autograd for the gradients)
Bm, Q = 4, 300
lg, bx, tg, off, tcat, oq, ot, on = run_match(Bm, Q, [50, 40, 0, 50], seed=8)
vr = (torch.rand(Bm, Q, 2) * 1.5 + 0.2).to(dev)
Kmax = oq.shape[1]
out6 = torch.empty(6, device=dev)
g_ce = torch.empty(Bm, Q, 2, device=dev); g_bbox = torch.empty(Bm, Q, 4, device=dev); g_giou = torch.empty(Bm, Q, 4, device=dev)
g_vb = torch.empty(Bm, Q, 4, device=dev); g_vv = torch.empty(Bm, Q, 2, device=dev); matched = torch.empty(Bm * Q, dtype=torch.uint8, device=dev)
nb = float(sum(len(t) for t in tg))
L.call("cdetr_set_loss_fwd", lg, bx, vr, tcat, off, oq, ot, on, Bm, Q, Kmax, torch.tensor([nb], device=dev), 1.0, 0.25, out6, g_ce, g_bbox, g_giou, g_vb, g_vv, matched, None)
lc, bc, vc = lg.cpu().requires_grad_(), bx.cpu().requires_grad_(), vr.cpu().requires_grad_()
targets = [{"boxes": t, "labels": torch.zeros(len(t), dtype=torch.int64)} for t in tg]
idx = [(oq[b, : on[b]].cpu(), ot[b, : on[b]].cpu()) for b in range(Bm)]
ol, _ = OC.set_criterion({"pred_logits": lc, "pred_boxes": bc, "pred_vars": vc}, targets, indices=idx)
names = ["loss_ce", "class_error", "loss_bbox", "loss_giou", "cardinality_error", "loss_variance"]
for i, n in enumerate(names):
    report(f"set_loss {n}", out6[i].cpu(), ol[n].detach(), 1e-5)
up = torch.tensor([2.0, 5.0, 2.0, 2.0], device=dev)
dl, dbx, dvr = torch.empty_like(lg), torch.empty_like(bx), torch.empty_like(vr)
L.call("cdetr_set_loss_bwd", up, g_ce, g_bbox, g_giou, g_vb, g_vv, Bm * Q, dl, dbx, dvr)
(2 * ol["loss_ce"] + 5 * ol["loss_bbox"] + 2 * ol["loss_giou"] + 2 * ol["loss_variance"]).backward()
report("set_loss d_logits", dl.cpu(), lc.grad, 2e-5); report("set_loss d_boxes", dbx.cpu(), bc.grad, 2e-5); report("set_loss d_vars", dvr.cpu(), vc.grad, 2e-5)
# stage-1 criterion
n = 600
pw = (torch.rand(n, 2) * 0.1 + 0.02); pts = torch.rand(n, 2) * 0.8 + 0.1; whs = torch.rand(n, 2) * 0.1 + 0.02
out2 = torch.empty(2, device=dev); gw = torch.empty(n, 2, device=dev); gg = torch.empty(n, 2, device=dev)
L.call("cdetr_bbox_loss_fwd", pw.to(dev), pts.to(dev), whs.to(dev), n, out2, gw, gg)
pwc = pw.clone().requires_grad_()
o1 = OC.bounding_box_criterion({"pred_wh": pwc[None]}, {"points": pts[None], "whs": whs[None]})
report("bbox_loss wh", out2[0].cpu(), o1["loss_wh"].detach(), 1e-5); report("bbox_loss giou", out2[1].cpu(), o1["loss_giou"].detach(), 1e-5)
(o1["loss_wh"] + 0.4 * o1["loss_giou"]).backward()
dwh = torch.empty(n, 2, device=dev)
L.call("cdetr_bbox_loss_bwd", torch.tensor([1.0, 0.4], device=dev), gw, gg, n, dwh)
report("bbox_loss d_wh", dwh.cpu(), pwc.grad, 2e-5)

# LSAP timing
for (Bm, Q, T) in [(16, 300, 50), (1, 1000, 1000), (148, 1000, 1000), (8, 500, 50)]:
    g = torch.Generator().manual_seed(0)
    cost = torch.rand(Bm, Q * T, generator=g).to(dev)
    off = torch.arange(0, (Bm + 1) * T, T, dtype=torch.int32, device=dev)
    K = min(Q, T)
    oq = torch.empty(Bm, K, dtype=torch.int64, device=dev); ot = oq.clone(); on = torch.zeros(Bm, dtype=torch.int32, device=dev); status = torch.zeros(1, dtype=torch.int32, device=dev)
    for _ in range(2): L.call("cdetr_lsap", cost, off, Bm, Q, T, oq, ot, on, status)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): L.call("cdetr_lsap", cost, off, Bm, Q, T, oq, ot, on, status)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"lsap time B={Bm} {Q}x{T}: {ms*1e3:.1f} us total, {ms*1e3/Bm:.1f} us/image", flush=True)

print("FAILS", len(fails), fails)
