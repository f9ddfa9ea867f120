"""CPU restatement of the reference's categorical-memory algorithm (TEST INFRASTRUCTURE).

This file is the parity oracle for ``pinthememory_b200``: a dtype-generic (fp32 / fp64),
device-agnostic, functional restatement in plain torch ops of what
``/root/reference/network/memory.py::Memory_sup`` computes, with autograd supplying the
reference gradients. It is *not* the product and the product never imports it.

Pinning: ``tests/test_oracle_golden.py`` checks every function here against
``tests/golden/*.npz`` (outputs of the unmodified reference, generated in the build
container by ``oracle/make_golden.py``) and, when ``/root/reference`` is present, against
the live reference module. The reference itself ships no tests or golden vectors
(SURVEY.md section 4 / 8c), so those fixtures are the only pin there is.

Every function cites the reference lines it restates (paths relative to /root/reference).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

IGNORE_LABEL = 255


# ----------------------------------------------------------------------------- read


def normalize_channels(x):
    """q = x / max(||x||_2, 1e-12) over the channel dim of an NCHW map (memory.py:215,319)."""
    return F.normalize(x, dim=1)


def similarity(x, memory):
    """Normalised NHWC query and its similarity to every memory slot.

    memory.py:319-320 (normalise, NCHW->NHWC) and memory.py:171 (query . memory^T).
    Returns (q [B,h,w,C], s [B,h,w,K]).
    """
    q = normalize_channels(x).permute(0, 2, 3, 1).contiguous()
    s = torch.matmul(q, memory.t())
    return q, s


def feature_cohesion_loss(s, labels, temperature):
    """Read loss: CE of the bilinearly up-sampled logits against full-res labels.

    memory.py:172-176: s/T -> [B,K,h,w] -> bilinear (align_corners=True) to the label
    size -> CrossEntropy(ignore_index=255), mean over non-ignored label pixels.
    """
    logits = (s / temperature).permute(0, 3, 1, 2).contiguous()
    logits = F.interpolate(logits, size=labels.shape[1:], mode="bilinear", align_corners=True)
    return F.cross_entropy(logits, labels, ignore_index=IGNORE_LABEL)


def gumbel_noise_pair(s_flat):
    """The two Gumbel draws in the reference's order (memory.py:183-184).

    ``F.gumbel_softmax`` draws ``-empty_like(logits).exponential_().log()``; the dim-0 call
    comes first, the dim-1 call second.
    """
    g_query = -torch.empty_like(s_flat).exponential_().log()
    g_memory = -torch.empty_like(s_flat).exponential_().log()
    return g_query, g_memory


def slot_scores(s_flat, noise=None):
    """score_query = softmax over all pixels (dim 0); score_memory = softmax over slots (dim 1).

    memory.py:181-187. ``noise`` = (g_query, g_memory) turns both into tau=1 soft
    Gumbel-softmax samples.
    """
    if noise is None:
        return F.softmax(s_flat, dim=0), F.softmax(s_flat, dim=1)
    g_query, g_memory = noise
    return F.softmax(s_flat + g_query, dim=0), F.softmax(s_flat + g_memory, dim=1)


def read(x, memory, labels=None, temperature=1.0, noise=None):
    """Memory read up to (not including) the 1x1 output conv. memory.py:317-336, 167-189.

    Returns dict(u [B,2C,h,w], score_query [B,h,w,K], score_memory [B,h,w,K], readloss, s).
    ``readloss`` is the python int 0 when ``labels`` is None (memory.py:178).
    """
    B, C, h, w = x.shape
    K = memory.shape[0]
    q, s = similarity(x, memory)
    readloss = feature_cohesion_loss(s, labels, temperature) if labels is not None else 0
    s_flat = s.reshape(B * h * w, K)
    score_query, score_memory = slot_scores(s_flat, noise)
    picked = torch.matmul(score_memory, memory)  # memory.py:328
    u = torch.cat((q.reshape(B * h * w, C), picked), dim=1)  # memory.py:330
    u = u.view(B, h, w, 2 * C).permute(0, 3, 1, 2).contiguous()  # memory.py:331-332
    return dict(u=u, score_query=score_query.view(B, h, w, K), score_memory=score_memory.view(B, h, w, K),
                readloss=readloss, s=s)


# ---------------------------------------------------------------------------- write


def soft_label_weights(labels, num_slots, h, w, dtype=torch.float32):
    """Bilinear soft one-hot weights at feature resolution. memory.py:220-225.

    255 -> slot K, one_hot(K+1), to float, bilinear down-sample (align_corners=True) to
    [h,w]. Returns omega [B, h*w, K+1]; rows sum to 1.
    """
    lab = labels.clone()
    lab[lab == IGNORE_LABEL] = num_slots
    onehot = F.one_hot(lab, num_classes=num_slots + 1).permute(0, 3, 1, 2).contiguous().type(dtype)
    omega = F.interpolate(onehot, [h, w], mode="bilinear", align_corners=True)
    return omega.permute(0, 2, 3, 1).contiguous().view(labels.shape[0], h * w, num_slots + 1)


def class_sums(f, labels, num_slots):
    """Per-class soft-masked sums of the normalised write feature and the soft counts.

    memory.py:215-231. Returns (S [K+1, C], D [K+1]); row/entry K is the ignore slot.
    """
    B, C, h, w = f.shape
    v = normalize_channels(f).view(B, C, h * w)
    omega = soft_label_weights(labels, num_slots, h, w, dtype=f.dtype)
    D = omega.sum(1).sum(0)
    S = torch.matmul(v, omega).sum(0).t()
    return S, D


def momentum_update(memory_old, S, D, momentum):
    """Per-slot momentum write then row re-normalisation. memory.py:233-239.

    Slots with a zero count keep their old row (then get re-normalised like the rest).
    ``memory_old`` is treated as a constant (memory.py:233 ``clone().detach()``; read()
    has already detached m_items at memory.py:323-324 when writing).
    """
    K = memory_old.shape[0]
    old = memory_old.detach()
    present = (D[:K] != 0).unsqueeze(1)
    safe = torch.where(D[:K] != 0, D[:K], torch.ones_like(D[:K])).unsqueeze(1)
    blended = momentum * old + (1.0 - momentum) * S[:K] / safe
    return F.normalize(torch.where(present, blended, old), dim=1)


def divergence_loss(memory):
    """Memory-divergence loss: mean positive off-diagonal cosine. memory.py:264-272."""
    K = memory.shape[0]
    gram = torch.matmul(memory, memory.t()).clamp_min(0)
    return (gram.sum() - torch.trace(gram)) / (K * (K - 1))


def classification_loss(memory, weight, bias):
    """CE of the slot classifier on the updated memory. memory.py:259-262."""
    target = torch.arange(memory.shape[0], device=memory.device)
    return F.cross_entropy(F.linear(memory, weight, bias), target, ignore_index=IGNORE_LABEL)


def write(f, labels, memory_old, momentum, cls_weight, cls_bias, reduce_fn=None):
    """Memory write after the writing net. memory.py:215-250.

    ``reduce_fn`` (optional) is applied to the packed [K+1, C+1] sums|counts before the
    update: the hook where a sharded run all-reduces.
    Returns dict(memory_new, div_loss, cls_loss, S, D).
    """
    K = memory_old.shape[0]
    S, D = class_sums(f, labels, K)
    if reduce_fn is not None:
        packed = reduce_fn(torch.cat((S, D.unsqueeze(1)), dim=1))
        S, D = packed[:, :-1], packed[:, -1]
    memory_new = momentum_update(memory_old, S, D, momentum)
    return dict(memory_new=memory_new, div_loss=divergence_loss(memory_new),
                cls_loss=classification_loss(memory_new, cls_weight, cls_bias), S=S, D=D)


def label_histogram(labels, num_slots):
    """Integer count of label pixels per class; ignore (255) goes to bin K."""
    lab = labels.reshape(-1).clone()
    lab[lab == IGNORE_LABEL] = num_slots
    return torch.bincount(lab, minlength=num_slots + 1)


# ------------------------------------------------------------------- module wrapper


def _init_like_reference(module):
    """memory.py:9-19 initialisation: conv kaiming-normal, BN (1, 1e-4), Linear N(0,1e-4)/0."""
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight.data, nonlinearity="relu")
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.data.fill_(1.0)
            m.bias.data.fill_(1e-4)
        elif isinstance(m, nn.Linear):
            m.weight.data.normal_(0.0, 0.0001)
            m.bias.data.zero_()


class ReluGates:
    """Tie-breaking for the two ReLUs of the module (memory.py:86,106) when the oracle checks ANOTHER implementation.

    A ReLU is discontinuous in its derivative: where a pre-activation z is zero to within rounding, two correct fp32
    implementations of the convolution + BatchNorm before it (cuDNN, MKL, 3xTF32 tensor cores: all ~1e-7..2e-6 apart)
    legitimately land on different sides, and everything downstream of that element's gradient then differs at O(1).
    On maps of 10^5..10^6 elements such ties are certain to occur. The checked implementation therefore records its
    gates (z > 0) per ReLU call; the oracle replays them in call order and REPORTS every element where its own gate
    differs, with |z| relative to max|z| -- a test accepts a replayed gate only if that ratio is at rounding level.
    """

    def __init__(self, gates=None):
        self.gates = {k: list(v) for k, v in (gates or {}).items()}  # name -> FIFO of bool tensors
        self.mismatches = []                                          # (name, count, numel, max |z| / max|z|)

    def apply(self, name, z):
        queue = self.gates.get(name)
        if not queue:
            return F.relu(z)
        gate = queue.pop(0).to(z.device)
        natural = z.detach() > 0
        diff = natural != gate
        n = int(diff.sum())
        if n:
            scale = float(z.detach().abs().max())
            self.mismatches.append((name, n, z.numel(), float(z.detach().abs()[diff].max()) / max(scale, 1e-30)))
        return z * gate.to(z.dtype)


class _WriteFeature(nn.Module):
    """relu(x + BN(conv1x1(x))). memory.py:67-87."""

    def __init__(self, dim):
        super().__init__()
        self.writefeat = nn.Sequential(nn.Conv2d(dim, dim, kernel_size=1, bias=False), nn.BatchNorm2d(dim))

    def forward(self, x, relu_gates=None):
        z = x + self.writefeat(x)
        return F.relu(z) if relu_gates is None else relu_gates.apply("writenet", z)


class OracleMemorySup(nn.Module):
    """Oracle with the reference module's interface, sub-module names and state_dict keys.

    Device-agnostic (the reference hard-codes .cuda()); used by the tests as the checker
    and by bench.py as the timed CPU baseline ("port").
    """

    def __init__(self, memory_size, input_feature_dim, feature_dim, momentum, temperature, gumbel_read):
        super().__init__()
        assert input_feature_dim == feature_dim
        self.memory_size, self.feature_dim = memory_size, feature_dim
        self.momentum, self.temperature, self.gumbel_read = momentum, temperature, gumbel_read
        self.output = nn.Sequential(nn.Conv2d(2 * feature_dim, input_feature_dim, kernel_size=1, bias=False),
                                    nn.BatchNorm2d(input_feature_dim), nn.ReLU(inplace=True))
        self.writenet = _WriteFeature(feature_dim)
        self.clsfier = nn.Linear(feature_dim, memory_size, bias=True)
        self.m_items = F.normalize(torch.rand((memory_size, feature_dim), dtype=torch.float), dim=1)
        _init_like_reference(self)
        self.reduce_fn = None
        self.relu_gates = None  # a ReluGates: replay another implementation's ReLU decisions (see its docstring)

    def forward(self, query, mask=None, memory_writing=True, writing_detach=True, noise=None):
        if memory_writing:  # memory.py:323-324
            self.m_items = self.m_items.detach()
        B, C, h, w = query.shape
        if self.gumbel_read and noise is None:
            with torch.no_grad():
                noise = gumbel_noise_pair(torch.empty(B * h * w, self.memory_size, dtype=query.dtype,
                                                      device=query.device))
        r = read(query, self.m_items, mask, self.temperature, noise if self.gumbel_read else None)
        if self.relu_gates is None:
            updated_query = self.output(r["u"])
        else:
            updated_query = self.relu_gates.apply("output", self.output[1](self.output[0](r["u"])))
        writeloss = [0, 0]
        if memory_writing:  # memory.py:199-200, 206-257
            f = self.writenet(query) if self.relu_gates is None else self.writenet(query, self.relu_gates)
            wr = write(f, mask, self.m_items, self.momentum, self.clsfier.weight,
                       self.clsfier.bias, self.reduce_fn)
            self.m_items = wr["memory_new"].detach() if writing_detach else wr["memory_new"]
            writeloss = [wr["div_loss"], wr["cls_loss"]]
        return updated_query, r["score_query"], r["score_memory"], r["readloss"], writeloss

    def get_score(self, query, mask, mem):
        """memory.py:167-189 on an already normalised NHWC query (called from train.py:894)."""
        B, h, w, _ = query.shape
        s = torch.matmul(query, mem.t())
        readloss = feature_cohesion_loss(s, mask, self.temperature) if mask is not None else 0
        s_flat = s.view(B * h * w, mem.shape[0])
        noise = gumbel_noise_pair(s_flat) if self.gumbel_read else None
        score_query, score_memory = slot_scores(s_flat, noise)
        return score_query, score_memory, readloss
