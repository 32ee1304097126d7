"""Average precision of the validation loop (ref: evaluations/detection.py:207-255).

The pair metrics of the same reference file (IoU, IoUConfidence, Orthogonity, MAE) are computed by the CUDA
kernel behind centerNetOffset.centerNetEvaluation (csrc/evaluate.cu).  What is left here is the aggregation over
the whole validation split, which the reference runs on the host over concatenated CPU lists
(trainer/model/centerOffsetRes10.py:18-105): a descending-score sweep and the reference's envelope integration.
The reference does both with Python loops over every detection (5760 tiles x up to 3000 pairs); here they are a
sort + cumulative sums, with identical results.
"""
import torch


def averagePrecisionPlots(ious, scores, objNum, threshold):
    """ref: detection.py:207-231.  Returns an (n, 2) float64 tensor of [recall, precision] rows (the reference
    returns the same numbers as a list of lists)."""
    ious = torch.as_tensor(ious).detach().cpu()
    scores = torch.as_tensor(scores).detach().cpu()
    if scores.numel() == 0:
        return torch.zeros(0, 2, dtype=torch.float64)
    order = torch.sort(scores)[1].flip(0)                     # the reference's tie order: ascending sort, flipped
    hit = (ious[order] >= threshold).to(torch.float64)
    tp = torch.cumsum(hit, 0)
    n = torch.arange(1, hit.numel() + 1, dtype=torch.float64)
    return torch.stack([tp / objNum, tp / n], 1)


def averagePrecisionAll(apPlots):
    """ref: detection.py:233-255.  Walking the plots from the last to the first, the reference keeps the running
    maximum y of the precision; whenever a strictly larger precision appears it closes the rectangle
    (x2 - x1) * y, where x2 is the recall at the previous maximum and x1 the recall of the entry just before the
    new one, and finally adds x2 * y.  Vectorised: the record positions are where the reversed running maximum
    increases."""
    p = torch.as_tensor(apPlots, dtype=torch.float64).reshape(-1, 2)
    if p.shape[0] == 0:
        return 0.0
    r = p.flip(0)
    recall, prec = r[:, 0], r[:, 1]
    prev_max = torch.cat([torch.zeros(1, dtype=torch.float64), torch.cummax(prec, 0)[0][:-1]])
    rec_pos = torch.nonzero(prec > prev_max).reshape(-1)     # entries that set a new maximum (y starts at 0)
    if rec_pos.numel() == 0:                                   # every precision is 0: ap = x2 * y = 1 * 0
        return 0.0
    ap = 0.0
    # rectangle closed by record j (j >= 1 in the list of records): y = precision at the previous record,
    # x2 = recall at the previous record, x1 = recall of the entry right before record j
    for a, b in zip(rec_pos[:-1].tolist(), rec_pos[1:].tolist()):
        ap += (float(recall[a]) - float(recall[b - 1])) * float(prec[a])
    last = int(rec_pos[-1])
    return ap + float(recall[last]) * float(prec[last])


def apAll(ious, scores, objNum, threshold):
    return averagePrecisionAll(averagePrecisionPlots(ious, scores, objNum, threshold))
