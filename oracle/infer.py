"""CPU restatement (numpy, as the reference does it) of the inference-side output formatting.

TEST INFRASTRUCTURE -- only tests/ may import this.  The reference code lives inside data-loader loops
(A2/infer.py:48-118, A1/engine.py:141-187) and cannot be imported as functions; the bodies are restated here line by
line (numpy float32 arrays scaled in place by integer image sizes, Python int() truncation), citing the lines.
"""
import numpy as np
import torch


def infer_select(out_logits, out_bbox, ref_points, ori_h, ori_w, threshold=0.5):
    """A2/infer.py:74-116 for ONE batch (the reference runs bs=1).  Returns the list of annotation dicts (without ids)."""
    prob = out_logits.sigmoid()                                   # :75
    object_prob = prob[..., 0]                                    # :76
    obj_pos = torch.where(object_prob >= threshold)               # :77-78
    pred_scores = object_prob[obj_pos[0], obj_pos[1]]             # :80
    pred_boxes = out_bbox[obj_pos[0], obj_pos[1]]                 # :81
    ref = ref_points[obj_pos[0], obj_pos[1]]                      # :82
    pred_points = ref.detach().cpu().numpy()                      # :83
    pred_points[..., 0] *= ori_w                                  # :84
    pred_points[..., 1] *= ori_h                                  # :85
    pred_boxes = pred_boxes.detach().cpu().numpy()                # :87
    pred_boxes[..., 0] *= ori_w                                   # :88-91
    pred_boxes[..., 1] *= ori_h
    pred_boxes[..., 2] *= ori_w
    pred_boxes[..., 3] *= ori_h
    annos = []
    for pred_score, pred_box, pred_point in zip(pred_scores, pred_boxes, pred_points):   # :102
        x_cen, y_cen, w, h = pred_box
        x_ref, y_ref = pred_point
        annos.append({"area": int(w * h), "bbox": [int(x_cen), int(y_cen), int(w), int(h)], "category_id": 1,
                      "score": float(pred_score), "point": [int(x_ref), int(y_ref)]})   # :105-113
    return annos, obj_pos[1].tolist()


def pseudo_label_format(points, pred_whs, orig_size):
    """A1/engine.py:148-166 for one image: points [1,Q,2], pred_whs [1,Q,2] tensors, orig_size integer pair."""
    points = points.squeeze(0).detach().cpu().numpy().copy()                  # :150
    pred_whs = torch.squeeze(pred_whs).detach().cpu().numpy().copy()          # :151
    orig_target_sizes = np.asarray(orig_size)                                 # :152
    pred_whs[:, 0] *= orig_target_sizes[0]                                    # :153-156
    pred_whs[:, 1] *= orig_target_sizes[1]
    points[:, 0] *= orig_target_sizes[0]
    points[:, 1] *= orig_target_sizes[1]
    annos = []
    for point, wh in zip(points, pred_whs):                                   # :157
        x_cen, y_cen = point
        w, h = wh
        annos.append({"area": int(w * h), "bbox": [int(x_cen), int(y_cen), int(w), int(h)], "category_id": 1,
                      "iscrowd": 0})                                          # :160-167
    return annos
