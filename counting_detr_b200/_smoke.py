"""__graft_entry__.smoke(): one tiny stage-2 train step on cuda:0, checked against the CPU oracle."""
import torch


def run():
    from . import synthetic as SY
    from .models import build_model
    from oracle import criterion as OC, model as OM       # the checker (test infrastructure)
    S, B, Q, T = 64, 1, 20, 4
    dev = torch.device("cuda", 0)
    model, crit, _ = build_model(SY.default_args(2, num_query_position=Q, device="cuda:0"))
    sd = SY.make_state_dict(SY.SynthCfg(stage=2, num_query_position=Q), 0)
    model.load_state_dict(sd, strict=True)
    model.to(dev).train()
    inp = SY.make_inputs(B, S, T=T, stage=2)
    out, _ = model(inp["image"].to(dev), None, inp["rects"])
    targets = [{k: v.to(dev) for k, v in t.items()} for t in inp["targets"]]
    ld = crit(out, targets)
    loss = sum(ld[k] * crit.weight_dict[k] for k in ld if k in crit.weight_dict)
    loss.backward()
    torch.cuda.synchronize()
    oo, _ = OM.forward(sd, OM.Config(stage=2, num_query_position=Q), inp["image"], rects=inp["rects"])
    ol, oidx = OC.set_criterion(oo, inp["targets"])
    gidx = crit.matcher(out, targets)
    for k in ("pred_logits", "pred_boxes", "pred_vars"):
        err = (out[k].detach().cpu() - oo[k]).abs().max().item()
        assert err <= 1e-3 * oo[k].abs().max().item(), (k, err)
    assert all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(gidx, oidx)), "matching differs"
    for k in ol:
        assert abs(ld[k].item() - ol[k].item()) <= 1e-3 * abs(ol[k].item()) + 1e-5, (k, ld[k].item(), ol[k].item())
    g = model.get_parameter("transformer.decoder_layers.5.ffn.norm2.weight").grad
    assert g is not None and torch.isfinite(g).all() and g.abs().sum() > 0
    print(f"smoke ok: loss={loss.item():.6f} oracle={sum(ol[k] * w for k, w in OC.STAGE2_WEIGHT_DICT.items()).item():.6f}")
