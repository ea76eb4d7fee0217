"""The empirical NTK through the functional savefield API.

Mirrors ``docs/examples/basic_usage/example_ntk_functorch.py:140-190`` of the reference: stack the two
data sets, feed them through the net and ``MSELoss(reduction="sum")``, accumulate
``param.vivit_ggn_exact["gram_mat"]()`` over the parameters in an extension hook, reorder and slice.
With that loss the factor of the loss Hessian is ``sqrt(2) I``, so the Gram matrix is twice the Jacobian
Gram: ``ntk[n, m, c, d] = <d f_c(x1_n) / d theta, d f_d(x2_m) / d theta>``.  Checked against autograd.
Kernels are replaced by the test double on this CPU-only machine.
"""

import pytest
import torch
from torch import nn

import tests._torch_kernels as double


@pytest.fixture(autouse=True)
def torch_kernels(monkeypatch):
    double.install(monkeypatch)


class AccumulateGramHook:
    def __init__(self, delete_buffers):
        self.gram = None
        self.delete_buffers = delete_buffers

    def __call__(self, module):
        for p in module.parameters(recurse=False):
            gram_p = p.vivit_ggn_exact["gram_mat"]()
            self.gram = gram_p if self.gram is None else self.gram + gram_p
            if self.delete_buffers:
                del p.vivit_ggn_exact


def empirical_ntk(net, x1, x2, delete_buffers=True):
    from vivit_b200 import ViViTGGNExact, backpack, extend

    n1 = x1.shape[0]
    X = torch.cat([x1, x2])
    net, loss_func = extend(net), extend(nn.MSELoss(reduction="sum"))
    hook = AccumulateGramHook(delete_buffers)
    with backpack(ViViTGGNExact(), extension_hook=hook):
        output = net(X)
        loss_func(output, torch.zeros_like(output)).backward()
    for p in net.parameters():
        p.grad = None
    gram = torch.einsum("cndm->nmcd", hook.gram)
    return 0.5 * gram[:n1, n1:]


def autograd_ntk(net, x1, x2):
    params = list(net.parameters())

    def jacobian(x):
        out = net(x)
        rows = []
        for n in range(out.shape[0]):
            for c in range(out.shape[1]):
                grads = torch.autograd.grad(out[n, c], params, retain_graph=True)
                rows.append(torch.cat([g.reshape(-1) for g in grads]))
        return torch.stack(rows).reshape(out.shape[0], out.shape[1], -1)

    j1, j2 = jacobian(x1), jacobian(x2)
    return torch.einsum("ncp,mdp->nmcd", j1, j2)


NETS = {
    "mlp": (lambda: nn.Sequential(nn.Linear(7, 6), nn.Tanh(), nn.Linear(6, 4), nn.Sigmoid(), nn.Linear(4, 3)),
            lambda n: torch.rand(n, 7)),
    "cnn": (lambda: nn.Sequential(nn.Conv2d(2, 3, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2), nn.Flatten(),
                                  nn.Linear(3 * 3 * 3, 2)),
            lambda n: torch.rand(n, 2, 6, 6)),
}


@pytest.mark.parametrize("delete_buffers", [True, False], ids=["free", "keep"])
@pytest.mark.parametrize("name", list(NETS), ids=list(NETS))
def test_empirical_ntk_matches_autograd(name, delete_buffers):
    torch.manual_seed(0)
    make_net, make_x = NETS[name]
    net = make_net().double()
    x1, x2 = make_x(3).double(), make_x(4).double()
    got = empirical_ntk(net, x1, x2, delete_buffers)
    want = autograd_ntk(net, x1, x2)
    assert got.shape == want.shape == (3, 4, want.shape[2], want.shape[2])
    assert torch.allclose(got, want, rtol=1e-9, atol=1e-12), (got - want).abs().max()
    for p in net.parameters():
        assert hasattr(p, "vivit_ggn_exact") != delete_buffers
        if not delete_buffers:
            del p.vivit_ggn_exact
