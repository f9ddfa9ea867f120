"""The three packed-label read-loss kernels (one thread per cell / per label row of a cell / a lane pair per cell) against
each other and against a float64 torch restatement of memory.py:173-176, on shapes that exercise their edge handling:
ragged maps, non-integer label/feature ratios, few slots, ignore labels, tiny temperatures (the exact per-pixel path) and
small grids (the row split of the fourth kernel)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [
    # B, h, w, Hm, Wm, K, T, ignore share
    (2, 24, 24, 192, 192, 19, 1.0, 0.1),     # ratio 8, the cfg-2 geometry
    (1, 12, 37, 97, 300, 19, 1.0, 0.2),      # ragged, non-integer ratios (8.7 / 8.3)
    (2, 16, 16, 256, 256, 19, 1.0, 0.0),     # ratio 17: small grid -> row split
    (1, 33, 18, 100, 55, 5, 0.5, 0.3),       # ratio 3.1 / 3.2, five slots
    (2, 20, 28, 160, 224, 19, 0.02, 0.1),    # tiny temperature: per-pixel maximum path
    (1, 9, 9, 72, 72, 1, 1.0, 0.5),          # one slot
    (1, 8, 40, 64, 320, 19, 1.0, 1.0),       # every label ignored: loss = 0/0 = NaN like torch, zero gradient
]


def _reference(s, labels, T, B, h, w, K):
    z = (s[:, :K].double().view(B, h, w, K).permute(0, 3, 1, 2) / T).requires_grad_(True)
    up = F.interpolate(z, size=labels.shape[1:], mode="bilinear", align_corners=True)
    lab = labels.clone()
    loss_sum = F.cross_entropy(up, lab, ignore_index=255, reduction="sum")
    V = int((lab != 255).sum())
    (g,) = torch.autograd.grad(loss_sum, z)
    # d(loss_sum)/d(s/T) in the kernels' [N, KP] layout
    return (loss_sum / V if V else torch.tensor(float("nan"))), g.permute(0, 2, 3, 1).reshape(B * h * w, K), V


@pytest.mark.parametrize("case", CASES)
def test_readloss_kernels_agree(case):
    from pinthememory_b200 import capi

    B, h, w, Hm, Wm, K, T, ign = case
    dev = "cuda"
    KP = capi.score_stride(K)
    g = torch.Generator(device="cpu").manual_seed(hash(case) % (2 ** 31))
    N = B * h * w
    s = torch.zeros(N, KP, device=dev)
    s[:, :K] = (torch.rand(N, K, generator=g) * 2 - 1).to(dev)
    labels = torch.randint(0, K, (B, Hm, Wm), generator=g)
    # piecewise-constant regions with some ignore
    blocks = torch.randint(0, K, (B, (Hm + 15) // 16, (Wm + 15) // 16), generator=g)
    labels = torch.where(torch.rand(B, Hm, Wm, generator=g) < 0.5,
                         blocks.repeat_interleave(16, 1).repeat_interleave(16, 2)[:, :Hm, :Wm], labels)
    labels[torch.rand(B, Hm, Wm, generator=g) < ign] = 255
    labels = labels.to(dev)
    ref_loss, ref_g, V = _reference(s, labels, T, B, h, w, K)
    outs = {}
    try:
        for gen in ("4", "3", "2"):
            os.environ["PINMEM_B200_READLOSS_GEN"] = gen
            buf = torch.zeros(N * KP + 2 * capi.WS_WORDS + 4, dtype=torch.float32, device=dev)
            ds, ws, out = buf[: N * KP], buf[N * KP: N * KP + 2 * capi.WS_WORDS], buf[N * KP + 2 * capi.WS_WORDS:]
            lab8 = capi.labels_pack(labels, K, ws)
            capi.readloss_fwd8(s, lab8, T, B, h, w, K, ds, ws, out)
            torch.cuda.synchronize()
            outs[gen] = (out[0].item(), ds.view(N, KP)[:, :K].double().clone(), ds.view(N, KP)[:, K:].abs().max().item())
    finally:
        os.environ.pop("PINMEM_B200_READLOSS_GEN", None)
    # the gradient is the difference of two sums of ~(Hm/h)*(Wm/w) bilinear weights per feature pixel (softmax part minus
    # one-hot part): with one slot it is exactly zero in the reference, so the error is measured against that natural scale
    area = (Hm / h) * (Wm / w)
    scale = max(ref_g.abs().max().item(), area)
    for gen, (loss, dsg, pad) in outs.items():
        assert pad == 0.0, f"generation {gen}: padded slots received gradient"
        if V == 0:
            assert loss != loss and dsg.abs().max().item() == 0.0
            continue
        assert abs(loss - ref_loss.item()) <= 1e-5 * max(1.0, abs(ref_loss.item())), (gen, loss, ref_loss.item())
        err = (dsg - ref_g).abs().max().item() / scale                # max-abs over max(|ref|, weights per feature pixel)
        rel = ((dsg - ref_g).norm() / max(ref_g.norm().item(), area)).item()   # relative L2
        assert err <= 1e-5 and rel <= 1e-5, (gen, err, rel)
