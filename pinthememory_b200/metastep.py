"""The memory module's part of one GS meta-train step of the reference (train.py:530-583), as a reusable sequence.

The reference runs four forwards and three backwards through the SAME memory module class per iteration:

  A  ``net(meta_train, memory_writing=True, writing_detach=False)`` + ``backward(retain_graph=True)``   (train.py:533-541)
     theta' = theta - inner_lr * grad, written into the leaf modules' ``_parameters`` of two model copies by
     ``put_theta`` (train.py:246-277): plain (non-leaf) tensors, so the outer gradient reaches theta through them
  B  ``updated_net2(meta_train, memory_writing=True, writing_detach=False)`` on the SAVED memory ``mem_t``
     (train.py:547-556): the new ``m_items`` carries graph to theta' of the writing net
  C  ``updated_net(meta_test, memory_writing=False)`` reading that graph-carrying memory + ``backward()``
     (train.py:558-575): gradients reach ``net``'s parameters through theta' and, for the writing net, through the memory
  D  ``net.eval(); net.m_items = mem_t; net(meta_train, memory_writing=True)`` under ``no_grad``   (train.py:578-583)

Only the memory module is in scope here (BASELINE config 3), so the "network" is the module itself
and the segmentation loss is replaced by ``<G, updated_query>``; loss weights are the reference's (train.py:1213-1215).
The same function drives ``pinthememory_b200.memory.Memory_sup`` (tests, bench) and the oracle (tests): it only uses
the public module interface.
"""
import torch

LOSS_W = dict(read=0.02, div=0.4, cls=0.2)


def put_theta(model, theta):
    """train.py:262-277: overwrite the ``_parameters`` of every leaf module with the tensors of ``theta``."""

    def rec(m, name=None):
        if len(m._modules) != 0:
            for k, v in m._modules.items():
                rec(v, str(k) if name is None else name + "." + k)
        else:
            for k, v in m._parameters.items():
                if isinstance(v, torch.Tensor):
                    m._parameters[k] = theta[name + "." + k]

    rec(model)
    return model


def updated_network(old, new, lr):
    """train.py:246-260 (load=False): theta' = theta - lr * grad for parameters that have a gradient."""
    params = dict(old.named_parameters())
    theta = {}
    for k, v in old.state_dict().items():
        if k in params and params[k].grad is not None:
            theta[k] = params[k] - lr * params[k].grad
        else:
            theta[k] = params[k] if k in params else v
    return put_theta(new, theta)


def _loss(out, G, with_write):
    uq, _, _, rl, wl = out
    loss = (uq * G.to(uq.dtype)).sum() + LOSS_W["read"] * rl
    if with_write:
        loss = loss + LOSS_W["div"] * wl[0] + LOSS_W["cls"] * wl[1]
    return loss


def meta_step(net, upd, upd2, x_tr, lab_tr, x_te, lab_te, G_tr, G_te, inner_lr=0.01):
    """One meta-train step on three instances of the same module class (``net`` holds the parameters, ``upd`` /
    ``upd2`` are the functional copies). Returns a dict of everything a parity test compares. ``net``'s ``.grad``
    fields hold the accumulated inner + outer gradients afterwards, ``net.m_items`` the final memory."""
    mem_t = net.m_items.clone().detach()                                   # train.py:530
    net.zero_grad(set_to_none=True)
    out_a = net(x_tr, lab_tr, memory_writing=True, writing_detach=False)    # A
    inner = _loss(out_a, G_tr, True)
    inner.backward(retain_graph=True)                                      # train.py:541
    inner_grads = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    updated_network(net, upd, inner_lr).train()
    updated_network(net, upd2, inner_lr).train()
    upd2.m_items = mem_t                                                   # train.py:547
    upd2(x_tr, lab_tr, memory_writing=True, writing_detach=False)           # B
    upd.m_items = upd2.m_items.clone()                                     # train.py:558
    out_c = upd(x_te, lab_te, memory_writing=False)                         # C
    outer = _loss(out_c, G_te, False)
    outer.backward()                                                       # train.py:574
    with torch.no_grad():                                                  # D, train.py:578-583
        net.eval()
        net.m_items = mem_t
        net(x_tr, lab_tr, memory_writing=True)
        net.train()
    return {"inner_loss": inner.detach(), "outer_loss": outer.detach(), "readloss_a": out_a[3].detach(),
            "div_a": out_a[4][0].detach(), "cls_a": out_a[4][1].detach(), "readloss_c": out_c[3].detach(),
            "uq_c": out_c[0].detach(), "memory_b": upd.m_items.detach(), "memory_final": net.m_items.detach(),
            "inner_grads": inner_grads,
            "grads": {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}}
