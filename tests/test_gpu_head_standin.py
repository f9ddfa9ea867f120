"""GPU: the module inside a segmentation head, called the way the reference's heads call it.

The reference tree is not on the GPU box, so the unchanged DeepLabV3+ head cannot run there (the container-side test,
tests/test_head_integration.py, builds it and checks that the call reaches this module). This test puts the module
between a stand-in "ASPP bottleneck" and a stand-in "decoder" with the reference's call sequence around the memory
(network/deepv3plus.py:555-580: bottleneck -> ``self.memory(dec0_up, gts, memory_writing, writing_detach)`` -> the first
output replaces the features -> classifier -> Upsample + CrossEntropy; the loss terms are weighted as train.py:1213-1215)
and compares losses and EVERY gradient -- head parameters on both sides of the module included -- with the same head around
the oracle module on the same GPU. The head's own layers are plain torch modules in both arms; only the memory differs.
"""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from gate_util import check_ties, oracle_module_like, record_gates
from golden_util import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _exact_fp32_convs():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


class Head(nn.Module):
    """bot_aspp -> memory -> final classifier, the memory slot of the DeepLabV3+ head (deepv3plus.py:555-580)."""

    def __init__(self, memory, c_in, c_mem, num_classes):
        super().__init__()
        self.bot_aspp = nn.Conv2d(c_in, c_mem, kernel_size=1, bias=False)
        self.memory = memory
        self.final = nn.Sequential(nn.Conv2d(c_mem, c_mem, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(c_mem),
                                   nn.ReLU(inplace=True), nn.Conv2d(c_mem, num_classes, kernel_size=1, bias=False))

    def forward(self, aspp, gts, memory_writing, writing_detach):
        dec0_up = self.bot_aspp(aspp)
        mem_output = self.memory(dec0_up, gts, memory_writing, writing_detach)   # deepv3plus.py:561
        dec0_up = mem_output[0]
        main_out = F.interpolate(self.final(dec0_up), size=gts.shape[-2:], mode="bilinear", align_corners=True)
        loss1 = F.cross_entropy(main_out, gts, ignore_index=255)
        return loss1, mem_output[3], mem_output[4]


def _total(loss1, readloss, writeloss):
    return loss1 + 0.02 * readloss + 0.4 * writeloss[0] + 0.2 * writeloss[1]   # train.py:1213-1215


@pytest.mark.parametrize("writing_detach", [True, False])
def test_head_around_the_module_matches_head_around_the_oracle(writing_detach):
    from pinthememory_b200 import synth
    from pinthememory_b200.memory import Memory_sup

    B, Cin, C, h, w, Hm, Wm, K = 2, 96, 64, 12, 16, 96, 128, 19
    torch.manual_seed(11)
    mem = Memory_sup(K, C, C, 0.8, 1.0, False).cuda()
    with torch.no_grad():
        mem.clsfier.weight.normal_(0, 0.2)
    mem.fold_min_pixels = 0   # the score-plane read + the tcgen05 GEMMs, as at full size
    head = Head(mem, Cin, C, K).cuda().train()
    aspp = synth.make_features(B, Cin, h, w, seed=41, device="cuda").requires_grad_(True)
    gts = synth.make_labels(B, Hm, Wm, K, "blocky", seed=42).cuda()
    state = {k: v.clone() for k, v in head.state_dict().items()}
    mem0 = mem.m_items.clone()

    with record_gates() as gates:
        total = _total(*head(aspp, gts, True, writing_detach))
        total.backward()
    got = {"loss": total.detach(), "d_aspp": aspp.grad.clone(), "memory": mem.m_items.detach().clone()}
    got.update({"grad " + n: p.grad.clone() for n, p in head.named_parameters()})

    mem.m_items = mem0
    head.load_state_dict(state)   # also rewinds the BatchNorm running statistics
    ora = oracle_module_like(mem, gates)
    head_o = Head(ora, Cin, C, K).cuda().train()
    head_o.load_state_dict(state)
    aspp_o = aspp.detach().clone().requires_grad_(True)
    total_o = _total(*head_o(aspp_o, gts, True, writing_detach))
    total_o.backward()
    check_ties(ora.relu_gates)

    assert_close(got["loss"], total_o.detach(), 1e-5, "total loss")
    assert_close(got["memory"], ora.m_items.detach(), 1e-5, "new memory")
    assert_close(got["d_aspp"], aspp_o.grad, 2e-5, "d loss / d ASPP features")
    for n, p in head_o.named_parameters():
        assert_close(got["grad " + n], p.grad, 2e-5, "grad " + n)
