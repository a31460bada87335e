"""Device output formats against the numpy restatement of the reference's host loops (-m gpu):
stage-2 inference selection (A2/infer.py:74-118), stage-1 pseudo-label records (A1/engine.py:148-166), PostProcess
top-k (A2/models/anchor_detr.py:370-402).  Integer fields bit-exact, scores to 1e-6."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, seed, lo=0.0, hi=1.0):
    from counting_detr_b200 import synthetic as SY
    return SY.uniform(f"fmt{seed}", shape, lo, hi, seed)


@pytest.mark.parametrize("Q,sizes", [(300, [(384, 512)]), (500, [(333, 517)]), (1100, [(800, 1216), (1023, 77)])])
def test_infer_select_matches_reference_loop(Q, sizes):
    from counting_detr_b200.infer import detections_to_annotations, select_detections
    from oracle import infer as OI
    B = len(sizes)
    logits = _rand((B, Q, 2), 1, -3.0, 3.0)
    logits[0, :7, 0] = torch.tensor([0.0, -1e-9, 1e-9, -3e-8, 5e-8, -1e-7, 1e-7])      # threshold edge: sigmoid == 0.5
    boxes = torch.cat([_rand((B, Q, 2), 2), _rand((B, Q, 2), 3, 0.01, 0.3)], -1)
    ref = _rand((B, Q, 2), 4)
    sel = select_detections({"pred_logits": logits.cuda(), "pred_boxes": boxes.cuda()}, ref.cuda(),
                            torch.tensor(sizes, dtype=torch.float32))
    got = detections_to_annotations(sel, list(range(B)))
    want, n = [], 0
    for b, (h, w) in enumerate(sizes):
        annos, qs = OI.infer_select(logits[b:b + 1], boxes[b:b + 1], ref[b:b + 1], np.int64(h), np.int64(w))
        assert sel["count"][b].item() == len(annos)
        assert sel["query"][b, : len(annos)].tolist() == qs
        want += [dict(a, image_id=b) for a in annos]
    assert len(got) == len(want) and len(want) > Q // 4
    for g, w_ in zip(got, want):
        assert g["bbox"] == w_["bbox"] and g["area"] == w_["area"] and g["point"] == w_["point"], (g, w_)
        assert g["image_id"] == w_["image_id"] and abs(g["score"] - w_["score"]) <= 1e-6


@pytest.mark.parametrize("Q,size", [(50, (640, 480)), (3, (1023, 767)), (777, (333, 517))])
def test_pseudo_labels_match_reference_loop(Q, size):
    from counting_detr_b200.infer import pseudo_labels, pseudo_to_annotations
    from oracle import infer as OI
    points = _rand((1, Q, 2), 5, 0.02, 0.98)
    whs = _rand((1, Q, 2), 6, 0.005, 0.4)
    bbox, area = pseudo_labels(points.cuda(), whs.cuda(), size)
    got = pseudo_to_annotations(bbox, area, image_id=1)
    want = OI.pseudo_label_format(points, whs, np.array(size))
    assert len(got) == Q
    for g, w_ in zip(got, want):
        assert g["bbox"] == w_["bbox"] and g["area"] == w_["area"] and g["iscrowd"] == 0


def test_postprocess_kernel_matches_oracle_and_torch_topk():
    from counting_detr_b200.models import PostProcess
    from oracle import criterion as OC
    for (B, Q) in [(2, 300), (1, 500), (3, 64)]:
        out = {"pred_logits": _rand((B, Q, 2), 7, -4.0, 4.0), "pred_boxes": _rand((B, Q, 4), 8)}
        sizes = torch.tensor([[480.0, 640.0], [512.0, 512.0], [333.0, 517.0]])[:B]
        r = PostProcess()({k: v.cuda() for k, v in out.items()}, sizes.cuda())
        o = OC.post_process(out, sizes)
        for a, b in zip(r, o):
            assert torch.allclose(a["scores"].cpu(), b["scores"], atol=1e-6)
            assert torch.equal(a["labels"].cpu(), b["labels"])
            assert torch.allclose(a["boxes"].cpu(), b["boxes"], atol=1e-4)
            assert (a["scores"][1:] <= a["scores"][:-1]).all()


def test_normalize_u8_is_bit_identical_to_totensor_normalize():
    """transforms.ToTensor() + Normalize (A2/data/fsc147.py:22-24,82; torchvision functional: to float32, div(255),
    sub_(mean), div_(std)) restated with plain torch ops on the CPU; the device kernel must match bit for bit."""
    from counting_detr_b200.data import IMAGENET_MEAN, IMAGENET_STD, DevicePrefetcher, normalize_u8
    g = torch.Generator().manual_seed(0)
    u8 = torch.randint(0, 256, (3, 96, 160, 3), generator=g, dtype=torch.uint8)
    ref = u8.permute(0, 3, 1, 2).contiguous().to(torch.float32).div(255)
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1); std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    ref = ref.sub_(mean).div_(std)
    for batch in DevicePrefetcher([{"image": u8}], "cuda"):
        got = normalize_u8(batch["image"])
    assert torch.equal(got.cpu(), ref)
