"""nets.py: the reference's network samplers around an arbitrary loss(net) callable (the CNN scripts' contract,
complex_nets/Mnist/CNN/PMP_CNN.py:20-51): proposals and acceptance on the device, forward passes with the caller."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cnn_and_loss():
    import torch
    import torch.nn.functional as F
    from torch import nn

    class Model(nn.Module):            # the layer pattern of PMP_CNN.py:20-43 at a size a test can afford
        def __init__(self):
            super().__init__()
            self.conv1 = nn.Conv2d(1, 4, 5)
            self.conv2 = nn.Conv2d(4, 6, 3)
            self.fc1 = nn.Linear(6 * 10 * 10, 32)
            self.fc2 = nn.Linear(32, 10)

        def forward(self, x):
            out = F.max_pool2d(F.relu(self.conv1(x)), 2, 2)
            out = F.relu(self.conv2(out)).view(x.size(0), -1)
            return F.log_softmax(self.fc2(F.relu(self.fc1(out))), dim=1)

    torch.manual_seed(0)
    g = torch.Generator().manual_seed(1)
    X = torch.randn(64, 1, 28, 28, generator=g)
    y = torch.randint(0, 10, (64,), generator=g)

    @torch.no_grad()
    def loss(net):                      # PMP_CNN.py:46-51
        return torch.nn.CrossEntropyLoss()(net(X), y) / 10
    return Model(), loss


@pytest.mark.parametrize("kind", ["PMP", "MP", "MH"])
def test_callable_loss_samplers(ctx, kind):
    import torch
    from oracle import oracle as o
    from pmp_mcmc_b200 import nets
    net, loss = _cnn_and_loss()
    cls = {"PMP": nets.PMPOptimizer, "MP": nets.MPOptimizer, "MH": nets.MetropolisOptimizer}[kind]
    opt = cls(net, 1e-3, loss, seed=11, ctx=ctx) if kind == "MH" else cls(net, 1e-3, loss, N=7, seed=11, ctx=ctx)
    assert opt.d == sum(p.numel() for p in net.parameters())
    theta0 = nets.flatten(net)
    # one device-proposed step: the proposals are the Philox tree about the current parameters, the log-targets are the callable's
    opt.step(0, uniforms=np.array([0.37]))
    P = 2 if kind == "MH" else 8
    tree, b, depth = (o.TREE_FLAT, P, 1) if kind != "PMP" else (o.TREE_BINARY, 2, 3)
    props = o.propose(tree, b, depth, opt.d, 1e-3, theta0, 11, 0)
    assert np.array_equal(ctx.read_proposals().view(np.uint32), props.view(np.uint32))
    losses = np.array([float(loss(nets.unflatten(props[p], like=net))) for p in range(P)])
    A_dev = ctx.read_logweights()
    if kind == "PMP":
        A = o.standardize(o.psp_logweights(-losses, props[:, :4].astype(np.float64), 3, use_kernel=False))
    elif kind == "MP":
        kt = o.mp_logweights(np.zeros(P), props.astype(np.float64) / np.sqrt(opt.d)) / P      # sum_k mean_dim logK / P up to a constant
        A = o.standardize(-losses + (kt - kt.mean()))
    if kind != "MH":
        np.testing.assert_allclose(A_dev, A, rtol=1e-6, atol=1e-6)
        nxt = o.draw_blocked(o.weights_from_log(A_dev), [0.37], "right")[0]
    else:
        nxt = 1 if 0.37 < np.exp(10000.0 * (losses[0] - losses[1])) else 0
    assert np.array_equal(nets.flatten(opt.net), props[nxt])
    assert opt.loss == pytest.approx(losses[nxt], rel=1e-6)
    # the reference's injection seam: caller-made proposal nets
    if kind != "MH":
        cand = [opt.net] + [opt.update(opt.net) for _ in range(7)]
        opt.step(1, cand, [torch.from_numpy(nets.flatten(n)) for n in cand], torch.tensor(opt.d), uniforms=np.array([0.81]))
        assert any(opt.net is n for n in cand)
    tr = opt.fit(num_steps=3)
    assert len(tr) >= 4 and np.all(np.isfinite(tr))
