"""CPU restatement of the matcher and the losses (TEST INFRASTRUCTURE, see oracle/model.py header).

Follows A2/models/matcher.py:197-247 (OriginalHungarianMatcher), A2/util/box_ops.py:17-67,
A2/models/segmentation.py:198-223 (sigmoid_focal_loss), A2/models/anchor_detr.py:166-367
(SetCriterion), A1/models/anchor_detr.py:317-337 (BoundingBoxCriterion), A2 :370-402 (PostProcess).
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _oracle_lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-s", "-C", _HERE])
        _LIB = C.CDLL(path)
        _LIB.oracle_lsap.restype = C.c_int
    return _LIB


# ------------------------------------------------------------------ boxes
def cxcywh_to_xyxy(b):
    cx, cy, w, h = b.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)


def pairwise_giou(a, b):
    """generalized_box_iou, box_ops.py:46-67 (a [N,4], b [M,4] xyxy) -> [N,M]."""
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.max(a[:, None, :2], b[None, :, :2])
    rb = torch.min(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    union = area_a[:, None] + area_b[None, :] - inter
    iou = inter / union
    lt2 = torch.min(a[:, None, :2], b[None, :, :2])
    rb2 = torch.max(a[:, None, 2:], b[None, :, 2:])
    wh2 = (rb2 - lt2).clamp(min=0)
    hull = wh2[..., 0] * wh2[..., 1]
    return iou - (hull - union) / hull


# ------------------------------------------------------------------ matcher
def match_cost(logits, boxes, tgt_boxes, w_class=2.0, w_bbox=5.0, w_giou=2.0):
    """Per-image cost block [Q,T] (matcher.py:221-242); labels are all 0 -> class-0 probability only.
    Summation order bbox, class, giou as in :242."""
    p = logits.sigmoid()[:, 0:1]
    neg = 0.75 * (p ** 2.0) * (-(1 - p + 1e-8).log())
    pos = 0.25 * ((1 - p) ** 2.0) * (-(p + 1e-8).log())
    c_class = (pos - neg).expand(-1, tgt_boxes.shape[0])
    c_bbox = torch.cdist(boxes, tgt_boxes, p=1)
    c_giou = -pairwise_giou(cxcywh_to_xyxy(boxes), cxcywh_to_xyxy(tgt_boxes))
    return w_bbox * c_bbox + w_class * c_class + w_giou * c_giou


def lsap_c(cost):
    """scipy.optimize.linear_sum_assignment restated in C (oracle/lsap.c). cost: 2-D array-like."""
    c = np.ascontiguousarray(np.asarray(cost, dtype=np.float64))
    nr, nc = c.shape
    k = min(nr, nc)
    a = np.empty(k, dtype=np.int64)
    b = np.empty(k, dtype=np.int64)
    rc = _oracle_lib().oracle_lsap(C.c_int(nr), C.c_int(nc), c.ctypes.data_as(C.c_void_p),
                                   a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise ValueError("cost matrix is infeasible")
    return a, b


def lsap_py(cost):
    """Same algorithm as oracle/lsap.c in pure Python (small cases only): the readable statement of
    scipy's shortest-augmenting-path solver, tie rule and output order."""
    c = np.asarray(cost, dtype=np.float64)
    transposed = c.shape[0] > c.shape[1]
    if transposed:
        c = c.T
    nr, nc = c.shape
    u, v = np.zeros(nr), np.zeros(nc)
    col4row = -np.ones(nr, dtype=np.int64)
    row4col = -np.ones(nc, dtype=np.int64)
    path = -np.ones(nc, dtype=np.int64)
    for cur in range(nr):
        remaining = list(range(nc - 1, -1, -1))
        spc = np.full(nc, np.inf)
        SR, SC = set(), set()
        min_val, i, sink = 0.0, cur, -1
        while sink == -1:
            SR.add(i)
            index, lowest = -1, np.inf
            for it, j in enumerate(remaining):
                r = min_val + c[i, j] - u[i] - v[j]
                if r < spc[j]:
                    path[j], spc[j] = i, r
                if spc[j] < lowest or (spc[j] == lowest and row4col[j] == -1):
                    lowest, index = spc[j], it
            min_val = lowest
            j = remaining[index]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            SC.add(j)
            remaining[index] = remaining[-1]
            remaining.pop()
        u[cur] += min_val
        for r in SR:
            if r != cur:
                u[r] += min_val - spc[col4row[r]]
        for j in SC:
            v[j] -= min_val - spc[j]
        j = sink
        while True:
            r = path[j]
            row4col[j] = r
            col4row[r], j = j, col4row[r]
            if r == cur:
                break
    if transposed:
        order = np.argsort(col4row)
        return col4row[order], order
    return np.arange(nr), col4row


def hungarian_match(outputs, targets, w_class=2.0, w_bbox=5.0, w_giou=2.0, solver=lsap_c):
    """OriginalHungarianMatcher.forward -> list of (idx_query int64[K], idx_target int64[K])."""
    res = []
    with torch.no_grad():
        for b, t in enumerate(targets):
            cost = match_cost(outputs["pred_logits"][b], outputs["pred_boxes"][b], t["boxes"], w_class, w_bbox, w_giou)
            i, j = solver(cost.numpy())
            res.append((torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)))
    return res


# ------------------------------------------------------------------ losses
def focal_loss(logits, onehot, num_boxes, alpha=0.25, gamma=2.0):
    """sigmoid_focal_loss, segmentation.py:198-223."""
    p = logits.sigmoid()
    ce = F.binary_cross_entropy_with_logits(logits, onehot, reduction="none")
    p_t = p * onehot + (1 - p) * (1 - onehot)
    loss = ce * (1 - p_t) ** gamma
    loss = (alpha * onehot + (1 - alpha) * (1 - onehot)) * loss
    return loss.mean(1).sum() / num_boxes


def set_criterion(outputs, targets, indices=None, num_classes=1, focal_alpha=0.25, world_size=1,
                  weights=(2.0, 5.0, 2.0), solver=lsap_c):
    """SetCriterion.forward (labels, boxes, cardinality, vars), anchor_detr.py:308-331."""
    logits, boxes, pvars = outputs["pred_logits"], outputs["pred_boxes"], outputs["pred_vars"]
    B, Q, Cn = logits.shape
    if indices is None:
        indices = hungarian_match(outputs, targets, *weights, solver=solver)
    num_boxes = max(float(sum(len(t["labels"]) for t in targets)) / world_size, 1.0)
    bidx = torch.cat([torch.full_like(s, i) for i, (s, _) in enumerate(indices)])
    sidx = torch.cat([s for s, _ in indices])
    # labels (:166-197): matched -> its label (0), unmatched -> num_classes; one-hot then drop last column
    tgt_cls_o = torch.cat([t["labels"][j] for t, (_, j) in zip(targets, indices)])
    tgt_cls = torch.full((B, Q), num_classes, dtype=torch.int64)
    tgt_cls[bidx, sidx] = tgt_cls_o
    onehot = torch.zeros(B, Q, Cn + 1)
    onehot.scatter_(2, tgt_cls[..., None], 1)
    onehot = onehot[..., :-1]
    losses = {"loss_ce": focal_loss(logits, onehot, num_boxes, focal_alpha) * Q}
    matched_logits = logits[bidx, sidx]
    if tgt_cls_o.numel() == 0:
        losses["class_error"] = torch.tensor(100.0)
    else:
        acc = (matched_logits.argmax(-1) == tgt_cls_o).float().sum() * (100.0 / tgt_cls_o.numel())
        losses["class_error"] = 100 - acc
    # boxes (:213-234)
    src = boxes[bidx, sidx]
    tgt = torch.cat([t["boxes"][j] for t, (_, j) in zip(targets, indices)], 0)
    losses["loss_bbox"] = (src - tgt).abs().sum() / num_boxes
    giou = torch.diag(pairwise_giou(cxcywh_to_xyxy(src), cxcywh_to_xyxy(tgt)))
    losses["loss_giou"] = (1 - giou).sum() / num_boxes
    # cardinality (:199-211), no grad
    with torch.no_grad():
        card = (logits.argmax(-1) != Cn - 1).sum(1).float()
        lens = torch.tensor([float(len(t["labels"])) for t in targets])
        losses["cardinality_error"] = (card - lens).abs().mean()
    # Laplace-style uncertainty on w/h (:264-289): the L1 term is a SCALAR mean over matched pairs
    sv = pvars[bidx, sidx]
    lw = (src[:, 2] - tgt[:, 2]).abs().mean() / sv[:, 0].abs() + sv[:, 0].log().abs()
    lh = (src[:, 3] - tgt[:, 3]).abs().mean() / sv[:, 1].abs() + sv[:, 1].log().abs()
    losses["loss_variance"] = ((lw + lh) / num_boxes).sum()
    return losses, indices


STAGE2_WEIGHT_DICT = {"loss_ce": 2.0, "loss_bbox": 5.0, "loss_giou": 2.0, "loss_variance": 2.0}
STAGE1_WEIGHT_DICT = {"loss_wh": 1.0, "loss_giou": 0.4}


def bounding_box_criterion(outputs, targets):
    """BoundingBoxCriterion.forward, A1/models/anchor_detr.py:317-337."""
    pts = targets["points"].flatten(0, 1)
    src_wh = outputs["pred_wh"].flatten(0, 1)
    tgt_wh = targets["whs"].flatten(0, 1)
    sb = torch.cat([pts, src_wh], -1)
    tb = torch.cat([pts, tgt_wh], -1)
    giou = torch.diag(pairwise_giou(cxcywh_to_xyxy(sb), cxcywh_to_xyxy(tb)))
    return {"loss_wh": (src_wh - tgt_wh).abs().mean(), "loss_giou": (1 - giou).sum() / tgt_wh.shape[0]}


def post_process(outputs, target_sizes, k=100):
    """PostProcess.forward, A2/models/anchor_detr.py:370-402."""
    logits, boxes = outputs["pred_logits"], outputs["pred_boxes"]
    B, Q, Cn = logits.shape
    prob = logits.sigmoid().view(B, -1)
    scores, idx = torch.topk(prob, k, dim=1)
    qi = idx // Cn
    labels = idx % Cn
    xyxy = torch.gather(cxcywh_to_xyxy(boxes), 1, qi[..., None].expand(-1, -1, 4))
    h, w = target_sizes.unbind(1)
    xyxy = xyxy * torch.stack([w, h, w, h], 1)[:, None, :]
    return [{"scores": s, "labels": l, "boxes": b} for s, l, b in zip(scores, labels, xyxy)]
