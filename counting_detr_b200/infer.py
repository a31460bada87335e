"""Inference-side output formats of the two-stage recipe on the device.

The reference formats predictions on the host after copying every tensor back (A2/infer.py:74-118 for stage 2,
A1/engine.py:141-187 for the stage-1 pseudo-label pass that feeds stage 2's dataset, A2/data/fsc147.py:18-19,81).
Here one kernel per step produces the integer records on the device; a single small D2H copy per image (or per epoch)
then yields exactly the dicts the reference appends to its COCO-style JSON.
"""
import torch

from . import _lib as L


@torch.no_grad()
def select_detections(outputs, ref_points, orig_sizes, threshold=0.5):
    """Stage-2 inference selection (A2/infer.py:74-118).  outputs: model output dict; ref_points [B,Q,2]; orig_sizes
    [B,2] = (ori_h, ori_w).  Returns device tensors: count [B] and, per image, the first count[b] rows of
    query [B,Q], score [B,Q], bbox [B,Q,4] (int cx,cy,w,h in pixels), area [B,Q], point [B,Q,2]."""
    logits = outputs["pred_logits"].detach().float().contiguous()
    boxes = outputs["pred_boxes"].detach().float().contiguous()
    dev = logits.device
    B, Q, C = logits.shape
    sizes = torch.as_tensor(orig_sizes, dtype=torch.float32).reshape(B, 2).to(dev).contiguous()
    res = {"count": torch.zeros(B, dtype=torch.int32, device=dev), "query": torch.zeros(B, Q, dtype=torch.int32, device=dev),
           "score": torch.zeros(B, Q, device=dev), "bbox": torch.zeros(B, Q, 4, dtype=torch.int32, device=dev),
           "area": torch.zeros(B, Q, dtype=torch.int32, device=dev), "point": torch.zeros(B, Q, 2, dtype=torch.int32, device=dev)}
    L.call("cdetr_infer_select", logits, C, boxes, ref_points.detach().float().contiguous(), sizes, B, Q, float(threshold),
           res["count"], res["query"], res["score"], res["bbox"], res["area"], res["point"])
    return res


def detections_to_annotations(sel, image_ids, first_anno_id=1):
    """The `predictions["annotations"]` records of A2/infer.py:99-116 (one D2H copy for the whole batch)."""
    host = {k: v.cpu() for k, v in sel.items()}
    out, anno_id = [], first_anno_id
    for b, image_id in enumerate(image_ids):
        for i in range(int(host["count"][b])):
            out.append({"id": anno_id, "image_id": int(image_id), "area": int(host["area"][b, i]),
                        "bbox": [int(v) for v in host["bbox"][b, i]], "category_id": 1,
                        "score": float(host["score"][b, i]), "point": [int(v) for v in host["point"][b, i]]})
            anno_id += 1
    return out


@torch.no_grad()
def pseudo_labels(points, pred_wh, orig_size):
    """Stage-1 pseudo-label records (A1/engine.py:148-166) for one image: points [Q,2] (or [1,Q,2]), pred_wh [Q,2] (or
    [1,Q,2]) normalised, orig_size = the reference's `orig_size` pair.  Returns device int32 bbox [Q,4] = [cx,cy,w,h] and
    area [Q]; `pseudo_to_annotations` turns them into the JSON records stage 2's dataset reads."""
    pts = points.detach().float().reshape(-1, 2).contiguous()
    wh = pred_wh.detach().float().reshape(-1, 2).contiguous()
    dev = wh.device
    n = wh.shape[0]
    size2 = torch.as_tensor(orig_size, dtype=torch.float32).reshape(-1)[:2].to(dev).contiguous()
    bbox = torch.zeros(n, 4, dtype=torch.int32, device=dev)
    area = torch.zeros(n, dtype=torch.int32, device=dev)
    L.call("cdetr_pseudo_label_format", pts.to(dev), wh, size2, n, bbox, area)
    return bbox, area


def pseudo_to_annotations(bbox, area, image_id, first_anno_id=1):
    b, a = bbox.cpu(), area.cpu()
    return [{"id": first_anno_id + i, "image_id": int(image_id), "area": int(a[i]), "bbox": [int(v) for v in b[i]],
             "category_id": 1, "iscrowd": 0} for i in range(b.shape[0])]
