"""
A second, independent witness for the oracle's restatement of the Keras arithmetic (no GPU needed): PyTorch's own CPU
``Linear`` / ``LSTM`` / autograd / ``Adam`` in float64.  TensorFlow/Keras cannot be installed here, so the oracle's Keras boundary
stays unpinned against the reference stack itself; this file rules out that the oracle's forward pass, its
hand-derived gradients (Dense with L1 activity terms; LSTM back-propagation through time) or its Adam step are wrong as
mathematics.  Conventions mapped: Keras kernels are ``[in, out]`` (torch: ``[out, in]``); both order the LSTM gates i, f, c/g, o;
Keras has one LSTM bias (torch: ``bias_ih`` + ``bias_hh``).
"""
import numpy as np
import pytest
import torch

from oracle import keras_math as km

ACT = {"tanh": torch.tanh, "sigmoid": torch.sigmoid, "relu": torch.relu, "linear": lambda z: z}


def t64(a, grad=False):
    return torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=grad)


def torch_ff(spec, weights, x, collect=None):
    a = x
    for (W, b), act in zip(weights, spec.acts):
        a = ACT[act](a @ W + b)
        if collect is not None:
            collect.append(a)
    return a


@pytest.mark.parametrize("dims,acts,l1", [((8, 7, 5, 4, 4, 5, 7, 8), None, None), ((6, 4, 6), ["relu", "linear"], [1e-3, 0.0]),
                                          ((5, 9, 3, 5), ["sigmoid", "tanh", "linear"], [0.0, 1e-2, 1e-4])])
def test_dense_forward_loss_and_gradients(dims, acts, l1):
    rng = np.random.default_rng(len(dims))
    spec = km.ff_hourglass_spec(dims[0]) if acts is None else km.FFSpec(list(dims), acts, l1)
    weights = [(W.astype(np.float64), rng.uniform(-0.3, 0.3, b.shape)) for W, b in km.init_ff_weights(spec, rng)]
    x, y = rng.random((12, spec.dims[0])), rng.random((12, spec.dims[-1]))
    loss, mse, grads, yhat = km.ff_loss_and_grads(spec, weights, x, y, dtype=np.float64)

    tw = [(t64(W, True), t64(b, True)) for W, b in weights]
    acts_t = []
    out = torch_ff(spec, tw, t64(x), acts_t)
    t_mse = ((out - t64(y)) ** 2).mean()
    t_loss = t_mse + sum(c * a.abs().sum() for c, a in zip(spec.l1, acts_t) if c)  # activity L1: coefficient * sum |a|, not divided by the batch
    t_loss.backward()
    np.testing.assert_allclose(yhat, out.detach().numpy(), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(km.ff_forward(spec, weights, x, np.float64), out.detach().numpy(), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose([loss, mse], [t_loss.item(), t_mse.item()], rtol=1e-12)
    for (gW, gb), (W, b) in zip(grads, tw):
        np.testing.assert_allclose(gW, W.grad.numpy(), rtol=1e-9, atol=1e-13)
        np.testing.assert_allclose(gb, b.grad.numpy(), rtol=1e-9, atol=1e-13)


def torch_lstm_stack(spec, weights, windows):
    layers, (Wd, bd) = weights
    seq = windows
    for (K, U, b), act in zip(layers, spec.acts):
        assert act == "tanh"  # torch.nn.LSTM is the tanh cell
        u = U.shape[0]
        cell = torch.nn.LSTM(K.shape[0], u, batch_first=True).double()  # (the global default dtype is left alone: other tests share the process)
        cell.weight_ih_l0, cell.weight_hh_l0 = torch.nn.Parameter(K.T.contiguous()), torch.nn.Parameter(U.T.contiguous())
        cell.bias_ih_l0, cell.bias_hh_l0 = torch.nn.Parameter(b.clone()), torch.nn.Parameter(torch.zeros_like(b))
        seq, _ = cell(seq)
    return ACT[spec.out_func](seq[:, -1] @ Wd + bd)


@pytest.mark.parametrize("units,lookback", [((6, 4, 6), 5), ((7,), 1), ((5, 3, 3, 5), 9)])
def test_lstm_forward_and_bptt_gradients(units, lookback):
    """The gate order, the single bias, 'last step of the last layer -> Dense', and BPTT through every layer and step."""
    rng = np.random.default_rng(sum(units))
    F = 4
    spec = km.LSTMSpec(n_features=F, units=list(units), acts=["tanh"] * len(units), n_features_out=F, out_func="linear", lookback_window=lookback)
    layers, dense = km.init_lstm_weights(spec, rng)
    layers = [(K.astype(np.float64), U.astype(np.float64), (b + rng.uniform(-0.2, 0.2, b.shape)).astype(np.float64)) for K, U, b in layers]
    dense = (dense[0].astype(np.float64), rng.uniform(-0.2, 0.2, dense[1].shape))
    weights = (layers, dense)
    windows, targets = rng.random((6, lookback, F)), rng.random((6, F))
    loss, grads, yhat = km.lstm_loss_and_grads(spec, weights, windows, targets, dtype=np.float64)

    tl = [(t64(K, True), t64(U, True), t64(b, True)) for K, U, b in layers]
    td = (t64(dense[0], True), t64(dense[1], True))
    out = torch_lstm_stack(spec, (tl, td), t64(windows))
    # torch.nn.Parameter() detaches: evaluate the stack functionally for autograd instead
    def functional(tl, td):
        seq = t64(windows)
        for (K, U, b) in tl:
            u = U.shape[0]
            h = torch.zeros(seq.shape[0], u, dtype=torch.float64)
            c = torch.zeros(seq.shape[0], u, dtype=torch.float64)
            outs = []
            for t in range(seq.shape[1]):
                z = seq[:, t] @ K + h @ U + b
                i, f, g, o = torch.sigmoid(z[:, :u]), torch.sigmoid(z[:, u:2 * u]), torch.tanh(z[:, 2 * u:3 * u]), torch.sigmoid(z[:, 3 * u:])
                c = f * c + i * g
                h = o * torch.tanh(c)
                outs.append(h)
            seq = torch.stack(outs, dim=1)
        return seq[:, -1] @ td[0] + td[1]

    out_f = functional(tl, td)
    np.testing.assert_allclose(out_f.detach().numpy(), out.detach().numpy(), rtol=1e-11, atol=1e-13)  # torch's fused LSTM == the spelled-out cell
    np.testing.assert_allclose(yhat, out.detach().numpy(), rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(km.lstm_forward_windows(spec, weights, windows, np.float64), out.detach().numpy(), rtol=1e-11, atol=1e-13)
    t_loss = ((out_f - t64(targets)) ** 2).mean()
    t_loss.backward()
    assert abs(loss - t_loss.item()) <= 1e-12 * abs(loss)
    (g_layers, (g_Wd, g_bd)) = grads
    for (gK, gU, gb), (K, U, b) in zip(g_layers, tl):
        np.testing.assert_allclose(gK, K.grad.numpy(), rtol=1e-8, atol=1e-13)
        np.testing.assert_allclose(gU, U.grad.numpy(), rtol=1e-8, atol=1e-13)
        np.testing.assert_allclose(gb, b.grad.numpy(), rtol=1e-8, atol=1e-13)
    np.testing.assert_allclose(g_Wd, td[0].grad.numpy(), rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(g_bd, td[1].grad.numpy(), rtol=1e-9, atol=1e-13)


def test_adam_steps_track_torch_adam():
    """
    Keras writes Adam as ``w -= lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps)``; torch as ``w -= lr/(1-b1^t) * m / (sqrt(v/(1-b2^t)) + eps)``.
    They differ only in where epsilon sits (by a factor sqrt(1-b2^t) on eps): with eps -> 0 the two are the same update, which pins the
    moment recursions and both bias corrections; with the default eps = 1e-7 the trajectories stay within that difference.
    """
    rng = np.random.default_rng(0)
    spec = km.FFSpec([5, 4, 5], ["tanh", "linear"])
    weights0 = [(W.astype(np.float64), rng.uniform(-0.1, 0.1, b.shape)) for W, b in km.init_ff_weights(spec, rng)]
    batches = [(rng.random((8, 5)), rng.random((8, 5))) for _ in range(25)]
    for eps, rtol in ((1e-30, 1e-10), (1e-7, 2e-3)):  # small gradients here (sqrt(v) ~ 1e-4): eps = 1e-7 is a 1e-3 effect, placed differently
        weights, state = [(W.copy(), b.copy()) for W, b in weights0], km.adam_init(weights0)
        tw = [(t64(W, True), t64(b, True)) for W, b in weights0]
        opt = torch.optim.Adam([p for pair in tw for p in pair], lr=1e-3, betas=(0.9, 0.999), eps=eps)
        for x, y in batches:
            _, _, grads, _ = km.ff_loss_and_grads(spec, weights, x, y, dtype=np.float64)
            weights = km.adam_step(weights, grads, state, lr=1e-3, b1=0.9, b2=0.999, eps=eps, dtype=np.float64)
            opt.zero_grad()
            ((torch_ff(spec, tw, t64(x)) - t64(y)) ** 2).mean().backward()
            opt.step()
        for (W, b), (tW, tb), (W0, b0) in zip(weights, tw, weights0):
            # compare the *movement* since initialisation: 25 steps of ~1e-3 each
            np.testing.assert_allclose(W - W0, tW.detach().numpy() - W0, rtol=rtol, atol=1e-12)
            np.testing.assert_allclose(b - b0, tb.detach().numpy() - b0, rtol=rtol, atol=1e-12)
