"""CPU restatement of the two callers either side of the memory path.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this.

* ``initial_prototypes`` -- ``Trainer.memory_initalize`` (reference train.py:1000-1042): pool the
  L2-normalised backbone features per class over the whole training set with the bilinearly
  down-sampled one-hot labels, divide by the soft counts (0 -> 1) and L2-normalise the rows.
* ``upsampled_cross_entropy`` -- the head's main loss (reference network/deepv3plus.py:575-578 with
  network/mynn.py:57-62 and loss.py:167-180): bilinear up-sampling (align_corners=True) of the class
  logits to label resolution, log-softmax over classes, mean NLL over the non-ignored pixels.

Pinned by tests/golden/callers/init_prototypes.npz and tests/golden/callers/main_loss.npz, which
oracle/make_golden_callers.py produced by executing the reference's own source for these functions
(the ``memory_initalize`` body extracted from /root/reference/train.py; ``Upsample`` and
``CrossEntropyLoss2d`` loaded from the reference files).
"""
import torch
import torch.nn.functional as F

from .memory_oracle import IGNORE_LABEL, normalize_channels, soft_label_weights


def pooled_class_sums(features, labels, num_slots):
    """One batch of train.py:1016-1034: (sums [K, C], counts [K]) of the normalised features."""
    B, C, h, w = features.shape
    q = normalize_channels(features).view(B, C, h * w)
    omega = soft_label_weights(labels, num_slots, h, w, dtype=features.dtype)  # [B, hw, K+1]
    counts = omega.sum(1)[:, :num_slots].sum(0)
    sums = torch.matmul(q, omega)[:, :, :num_slots].sum(0).t()
    return sums, counts


def initial_prototypes(batches, num_slots):
    """train.py:1003-1040 over an iterable of (features [B,C,h,w], labels [B,Hm,Wm]); returns
    (memory [K,C], sums [K,C], counts [K])."""
    sums = counts = None
    for features, labels in batches:
        s, c = pooled_class_sums(features, labels, num_slots)
        sums = s if sums is None else sums + s
        counts = c if counts is None else counts + c
    safe = torch.where(counts == 0, torch.ones_like(counts), counts)
    return F.normalize(sums / safe.unsqueeze(1), dim=1), sums, counts


def upsampled_cross_entropy(logits, labels):
    """mynn.py:57-62 + loss.py:175-180: mean over non-ignored label pixels (NaN when there is none)."""
    up = F.interpolate(logits, size=labels.shape[-2:], mode="bilinear", align_corners=True)
    return F.nll_loss(F.log_softmax(up, dim=1), labels, reduction="mean", ignore_index=IGNORE_LABEL)


def class_mean_vectors(features, labels, num_class):
    """tsnelib.py:48-74 (``input2basket``): normalise, bilinear up-sample (align_corners=True) to the label size, per-class
    sum over the hard one-hot labels (255 -> extra class), divide by the pixel count. Returns (means [K,C], counts [K]);
    rows of empty classes are zero (the reference skips them)."""
    b, c, h, w = features.shape
    H, W = labels.shape[-2:]
    f = F.interpolate(F.normalize(features, dim=1), [H, W], mode="bilinear", align_corners=True).view(b, c, -1)
    gt = labels.clone()
    gt[gt == IGNORE_LABEL] = num_class
    onehot = F.one_hot(gt, num_classes=num_class + 1).view(b, -1, num_class + 1).to(features.dtype)
    counts = onehot.sum(1).sum(0)[:num_class]
    sums = torch.matmul(f, onehot).sum(0).t()[:num_class]
    safe = torch.where(counts == 0, torch.ones_like(counts), counts)
    return sums / safe.unsqueeze(1), counts
