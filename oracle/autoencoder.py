"""Restatement of AutoEncoderModel (code/literal_encoder.py:19-144) with torch on CPU (float64).

Affine encoder/decoder stacks (shipped config: "encoder_active": "thah" matches neither branch at
:75-78 -> no activation), optional GLOBAL l2_normalize of the code (:66),
loss = mean((decoder - batch)^2) (:68), one AdagradOptimizer over all weights and biases
(acc0 = 0.1, no epsilon [TF semantics]); weights AND biases ~ N(0, 1) (tf.random_normal_initializer).
Parameter order (also the order of the flat vector of multike_b200/refapi/literal_encoder.py):
encoder_h0, encoder_b0, ..., decoder_h0, decoder_b0, ...
"""
import numpy as np
import torch

from .tf_semantics import ADAGRAD_INIT, l2_normalize


def shapes(input_dimension, hidden_dimensions):
    hds = [input_dimension] + list(hidden_dimensions)
    n = len(hidden_dimensions)
    out = []
    for i in range(n):
        out += [(hds[i], hds[i + 1]), (hds[i + 1],)]
    for i in range(n):
        j = n - i
        out += [(hds[j], hds[j - 1]), (hds[j - 1],)]
    return out


def activation(x, kind):
    if kind == "sigmoid":
        return torch.sigmoid(x)
    if kind == "tanh":
        return torch.tanh(x)
    return x


def forward(params, batch, n_layers, active, normalize):
    x = batch
    for i in range(n_layers):
        x = activation(x @ params[2 * i] + params[2 * i + 1], active)
    code = l2_normalize(x) if normalize else x
    y = code
    for i in range(n_layers):
        y = activation(y @ params[2 * (n_layers + i)] + params[2 * (n_layers + i) + 1], active)
    return x, y


class AutoEncoderOracle:
    def __init__(self, init_params, n_layers, active="thah", normalize=True, lr=0.001):
        self.params = [torch.as_tensor(np.asarray(p), dtype=torch.float64).clone() for p in init_params]
        self.acc = [torch.full_like(p, ADAGRAD_INIT) for p in self.params]
        self.n_layers, self.active, self.normalize, self.lr = n_layers, active, normalize, lr

    def step(self, batch):
        batch = torch.as_tensor(np.asarray(batch), dtype=torch.float64)
        ps = [p.clone().requires_grad_(True) for p in self.params]
        _, y = forward(ps, batch, self.n_layers, self.active, self.normalize)
        loss = ((y - batch) ** 2).mean()
        grads = torch.autograd.grad(loss, ps)
        for p, a, g in zip(self.params, self.acc, grads):
            a += g * g
            p -= self.lr * g * torch.rsqrt(a)
        return float(loss)

    def encode(self, data):
        x = torch.as_tensor(np.asarray(data), dtype=torch.float64)
        for i in range(self.n_layers):
            x = activation(x @ self.params[2 * i] + self.params[2 * i + 1], self.active)
        return x.numpy()
