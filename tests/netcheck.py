"""Net-level comparison of the CUDA engine with the CPU oracle on a common activation pattern (used by the -m gpu suites
and mirrored by __graft_entry__.smoke())."""
import copy

import torch

import srb200
from srb200 import models as M
from srb200 import functional as F
from oracle import torch_ref as R
from util import loss_of, rel_l2

DEV = "cuda:0"


def run_against_oracle(name, args, loss_kind, x, tgt, math, ref=None):
    """Returns dict(y_err, grad_errs{name: err}, n_convs, y, y_ref).  `ref`: an oracle net (built when None)."""
    if ref is None:
        ref = R.build(name, args, seed=0)
    ref.train()
    srb200.set_math(math)
    net = M.MODELS[name](*args)
    net.load_state_dict(ref.state_dict())
    net.to(DEV).train()
    masks = []
    F.record_activation_masks(masks)
    l0 = srb200.launch_count()
    y = net(x.to(DEV))
    F.record_activation_masks(None)
    assert srb200.launch_count() > l0, "the CUDA engine did not run"
    # dL/dy evaluated at the GPU's output, used by both backward passes (L1's sign(y - t) is discontinuous too)
    yc = y.detach().cpu().requires_grad_(True)
    loss_of(loss_kind, yc, tgt).backward()
    dy = yc.grad
    y.backward(dy.to(DEV))
    torch.cuda.synchronize()
    y_ref = ref(x).detach()
    forced = R.with_forced_activations(copy.deepcopy(ref), [m.cpu() for m in masks])
    forced.train()
    y_f = forced(x)
    assert not forced._forced_feed, "activation count mismatch between the engine and the oracle"
    y_f.backward(dy)
    gscale = max(p.grad.abs().max().item() for p in forced.parameters())
    errs = {}
    for (k, p), (_, q) in zip(net.named_parameters(), forced.named_parameters()):
        if q.grad.abs().max().item() < 1e-5 * gscale:
            continue  # mathematically-zero gradients (conv bias in front of BatchNorm): rounding noise only
        if q.numel() == 1:
            # one shared PReLU slope: sum(dy*z*[z<=0]) cancels to ~1e-4 of its terms; gate it against the gradient
            # scale of the net rather than against itself
            errs[k] = abs(p.grad.item() - q.grad.item()) / gscale
        else:
            errs[k] = rel_l2(p.grad, q.grad)
    return {"y_err": rel_l2(y.detach(), y_ref), "y_forced_err": rel_l2(y.detach(), y_f.detach()), "grad_errs": errs,
            "n_convs": R.count_convs(ref), "y": y.detach().cpu(), "y_ref": y_ref, "ref": ref}


def tolerances(math, n_convs):
    if math == "exact" or math == "fp32":
        return 1e-3, 1e-3
    ty = max(1e-3, 4.5e-4 * n_convs ** 0.5)
    return ty, 2 * ty


