"""Mirror of AutoEncoderModel (code/literal_encoder.py:19-144): the literal auto-encoder
1500 -> 1024 -> 512 -> dim -> 512 -> 1024 -> 1500, the only true GEMMs of the pipeline.

Same constructor and methods as the reference.  The six affine layers are plain dense GEMMs: forward, input
gradients and weight gradients (17 products per step) run on the hand-written tcgen05 / TMA / TMEM kernel of
csrc/mke_gemm.cu (multike_b200/gemm.py) at fp32-equivalent precision (3xTF32 split, chunked accumulation); the
chain rule between them (activation, global l2-norm of the code, squared error) is written out below -- no
autograd.  `args.encoder_gemm = "cublas"` runs the same step on torch.matmul in full fp32 instead: the timed
baseline of tools/bench_autoencoder.py.  All weights and biases live in ONE flat fp32 CUDA vector whose Adagrad
update is the hand-written dense kernel mke_dense_apply_adagrad (acc0 = 0.1, no epsilon [TF semantics]).
No CPU path.
"""
import sys as _sys

if __name__ == "literal_encoder":  # imported under the reference's top-level name (refapi first on sys.path):
    import multike_b200.refapi.literal_encoder as _canonical  # one module object, whichever name imported it first
    _sys.modules[__name__] = _canonical
import time

import numpy as np
import torch

from multike_b200 import _cabi
from multike_b200 import gemm as _gemm
from multike_b200.tables import ADAGRAD_INIT


def _shapes(input_dimension, hidden_dimensions):
    hds = [input_dimension] + list(hidden_dimensions)
    n = len(hidden_dimensions)
    out = []
    for i in range(n):
        out += [(hds[i], hds[i + 1]), (hds[i + 1],)]      # encoder_h{i}, encoder_b{i}   (:45-50)
    for i in range(n):
        j = n - i
        out += [(hds[j], hds[j - 1]), (hds[j - 1],)]      # decoder_h{i}, decoder_b{i}   (:51-60)
    return out


def _act(x, kind):
    if kind == 'sigmoid':
        return torch.sigmoid(x)
    if kind == 'tanh':
        return torch.tanh(x)
    return x  # the shipped "thah" matches neither branch (:75-78)


class AutoEncoderModel:
    def __init__(self, word_vec_list, args, input_dimension=1500, hidden_dimensions=None, init_params=None,
                 generator=None):
        self._lib = _cabi.load()
        self.session = None
        self.args = args
        self.device = torch.device(getattr(args, "device", "cuda"))
        self.input_dimension = input_dimension
        if hidden_dimensions is None:
            hidden_dimensions = [1024, 512, self.args.dim]
        self.hidden_dimensions = list(hidden_dimensions)
        self.layer_num = len(self.hidden_dimensions)
        self.gemm = str(getattr(args, "encoder_gemm", "tcgen05"))   # "tcgen05" (ours) | "cublas" (fp32 baseline)
        assert self.gemm in ("tcgen05", "cublas")
        data = torch.as_tensor(np.reshape(np.asarray(word_vec_list, dtype=np.float32),
                                          [len(word_vec_list), input_dimension])).to(self.device)
        if self.args.encoder_normalize:  # sklearn.preprocessing.normalize: unit l2 rows (:34-35)
            data = data / torch.clamp(data.norm(dim=1, keepdim=True), min=1e-12)
        self.word_vec_list = data
        shapes = _shapes(input_dimension, self.hidden_dimensions)
        total = sum(int(np.prod(s)) for s in shapes)
        self.theta = torch.empty(total, dtype=torch.float32, device=self.device)
        if init_params is None:  # tf.random_normal_initializer: N(0, 1) for weights AND biases
            self.theta.copy_(torch.randn(total, generator=generator).to(self.device))
        self.grad = torch.zeros_like(self.theta)
        self.acc = torch.full_like(self.theta, ADAGRAD_INIT)
        self.params, off = [], 0
        for k, s in enumerate(shapes):
            n = int(np.prod(s))
            view = self.theta[off:off + n].view(*s)
            if init_params is not None:
                view.copy_(torch.as_tensor(np.asarray(init_params[k], dtype=np.float32)).to(self.device))
            self.params.append(view)
            off += n
        self.weights = {('encoder_h%d' % i): self.params[2 * i] for i in range(self.layer_num)}
        self.weights.update({('decoder_h%d' % i): self.params[2 * (self.layer_num + i)] for i in range(self.layer_num)})
        self.biases = {('encoder_b%d' % i): self.params[2 * i + 1] for i in range(self.layer_num)}
        self.biases.update({('decoder_b%d' % i): self.params[2 * (self.layer_num + i) + 1] for i in range(self.layer_num)})

    # -- graph ------------------------------------------------------------------------------
    def encoder(self, input_data, params=None):
        params = self.params if params is None else params
        x = input_data
        for i in range(self.layer_num):
            x = _act(x @ params[2 * i] + params[2 * i + 1], self.args.encoder_active)
        return x

    def decoder(self, input_data, params=None):
        params = self.params if params is None else params
        x = input_data
        for i in range(self.layer_num):
            x = _act(x @ params[2 * (self.layer_num + i)] + params[2 * (self.layer_num + i) + 1], self.args.encoder_active)
        return x

    # -- one training step, chain rule written out ------------------------------------------------
    def _nt(self, a, b, bias=None):
        """a [M, K] . b [N, K]^T (+ bias): the tensor-core kernel, or fp32 cuBLAS as the baseline"""
        if self.gemm == "tcgen05":
            return _gemm.gemm_nt(a, b, bias)
        out = a @ b.t()
        return out if bias is None else out + bias

    def _act_grad(self, y, g):
        kind = self.args.encoder_active
        if kind == 'sigmoid':
            return g * y * (1 - y)
        if kind == 'tanh':
            return g * (1 - y * y)
        return g

    def _step(self, batch):
        """one session.run([loss, optimizer]) (:62-69): mean squared reconstruction error, Adagrad"""
        prev = torch.backends.cuda.matmul.allow_tf32
        if self.gemm == "cublas":
            torch.backends.cuda.matmul.allow_tf32 = False
        try:
            n_l, P = self.layer_num, self.params
            # forward: layers 0 .. L-1 encode, L .. 2L-1 decode; acts[i] = input of layer i
            acts, x, code, inv = [batch], batch, None, None
            for i in range(2 * n_l):
                x = _act(self._nt(x, P[2 * i].t().contiguous(), P[2 * i + 1]), self.args.encoder_active)
                if i == n_l - 1 and self.args.encoder_normalize:  # tf.nn.l2_normalize without axis: global norm (:66)
                    code = x
                    inv = torch.rsqrt(torch.clamp((code * code).sum(), min=1e-12))
                    x = code * inv
                acts.append(x)
            diff = acts[-1] - batch
            loss = (diff * diff).mean()
            # backward
            g = diff * (2.0 / diff.numel())
            off_end = self.theta.numel()
            for i in reversed(range(2 * n_l)):
                if i == n_l - 1 and code is not None:
                    # y = c r, r = rsqrt(max(S, eps)), S = sum c^2: dc = g r - c r^3 (g . c) while S >= eps
                    dot = (g * code).sum()
                    below = (code * code).sum() < 1e-12
                    g = g * inv - torch.where(below, torch.zeros_like(dot), dot * inv ** 3) * code
                    g = self._act_grad(code, g)
                else:
                    g = self._act_grad(acts[i + 1], g)
                w, b = P[2 * i], P[2 * i + 1]
                off_b = off_end - b.numel()
                off_w = off_b - w.numel()
                self.grad[off_b:off_end].copy_(g.sum(0))
                # dW [in, out] = X^T [in, B] . (dY^T [out, B])^T
                dw = self._nt(acts[i].t().contiguous(), g.t().contiguous())
                self.grad[off_w:off_b].copy_(dw.reshape(-1))
                if i > 0:
                    g = self._nt(g, w)   # dX [B, in] = dY [B, out] . (W [in, out])^T
                off_end = off_w
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        _cabi.check(self._lib.mke_dense_apply_adagrad(self.theta.data_ptr(), self.grad.data_ptr(), self.acc.data_ptr(),
                                                      self.theta.numel(), float(self.args.learning_rate),
                                                      _cabi.current_stream()))
        return loss.detach()

    def train_one_epoch(self, epoch):
        start_time = time.time()
        batch_size = self.args.batch_size
        num_batch = len(self.word_vec_list) // batch_size + 1  # (:96) the last batch may be empty: skipped here
        loss_sum = torch.zeros((), dtype=torch.float32, device=self.device)
        for i in range(num_batch):
            batch = self.word_vec_list[i * batch_size:(i + 1) * batch_size]
            if batch.shape[0] == 0:
                continue
            loss_sum += self._step(batch)
        loss_sum = float(loss_sum) + self.args.batch_size  # sic (:106)
        print('epoch {} of literal encoder, loss: {:.4f}, time: {:.4f}s'.format(epoch, loss_sum, time.time() - start_time))
        return loss_sum

    def encoder_multi_batches(self, input_data):
        """:114-144: the final encoding, float64, on the raw (un-normalised) inputs"""
        print('encode literal embeddings...', len(input_data))
        x = torch.as_tensor(np.reshape(np.asarray(input_data, dtype=np.float64), [len(input_data), self.input_dimension]))
        out = np.zeros((len(input_data), self.hidden_dimensions[-1]))
        batch_size = self.args.batch_size
        params = [p.double() for p in self.params[:2 * self.layer_num]]
        for a in range(0, len(input_data), batch_size):
            h = x[a:a + batch_size].to(self.device)
            for i in range(self.layer_num):
                h = _act(h @ params[2 * i] + params[2 * i + 1], self.args.encoder_active)
            out[a:a + batch_size] = h.cpu().numpy()
        print("encoded literal embeddings", out.shape)
        return out


# ---- the literal pipeline around the auto-encoder (code/literal_encoder.py:147-180) -----------------
def generate_unlisted_word2vec(word2vec, literal_list):
    """words of the literals that the word-vector table lacks get character-level vectors
    (utils.generate_word2vec_by_character_embedding: a gensim Word2Vec over characters -- host code of
    the reference, imported from its `utils` when there is anything to embed)"""
    missing = [w for literal in literal_list for w in literal.split(' ') if w not in word2vec]
    if missing:
        from utils import generate_word2vec_by_character_embedding  # the reference's module (needs gensim)
        word2vec.update(generate_word2vec_by_character_embedding(missing))
    return word2vec


def literal_token_matrix(literal_list, word2vec, tokens_max_len=5, word2vec_dimension=300):
    """[L, tokens_max_len, dim] float32: the vectors of each literal's first tokens, zeros elsewhere"""
    out = np.zeros((len(literal_list), tokens_max_len, word2vec_dimension), dtype=np.float32)
    for row, literal in zip(out, literal_list):
        for i, word in enumerate(literal.split(' ')[:tokens_max_len]):
            vec = word2vec.get(word)
            if vec is not None:
                row[i] = vec
    return out


class LiteralEncoder:
    """literals -> token vectors -> auto-encoder trained for args.encoder_epoch epochs -> codes
    (`encoded_literal_vector`, float64 [L, args.dim]); what data_model.py:81-83 instantiates"""

    def __init__(self, literal_list, word2vec, args, tokens_max_len=5, word2vec_dimension=300):
        self.args = args
        self.literal_list = literal_list
        self.word2vec = generate_unlisted_word2vec(word2vec, literal_list)
        self.tokens_max_len = tokens_max_len
        self.word2vec_dimension = word2vec_dimension
        tokens = literal_token_matrix(literal_list, self.word2vec, tokens_max_len, word2vec_dimension)
        model = AutoEncoderModel(tokens, args, input_dimension=tokens_max_len * word2vec_dimension)
        for epoch in range(1, args.encoder_epoch + 1):
            model.train_one_epoch(epoch)
        self.encoded_literal_vector = model.encoder_multi_batches(tokens)
