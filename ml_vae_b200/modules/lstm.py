"""Drop-in for ``torch.nn.LSTM`` as the reference's recipes declare it in yaml -- ``rnn: !new:torch.nn.LSTM`` with input_size,
hidden_size, num_layers, batch_first: True, dropout (src/models/MD_VAE/model.yaml:78-83 and the other MD_VAE* / test_h_vae
recipes), and as modules/boundary_detector.py:19 / phoneme_recognizer.py:13 / decoder.py:14-15 construct it: same constructor
keywords, same parameter names (weight_ih_l{k}[_reverse], ...: checkpoints load either way), same ``(output, (h_n, c_n))`` return.

bf16 CUDA activations run on the persistent tcgen05 recurrence (csrc/lstm.cu; one or two directions) with the time-parallel
products on the TMA GEMM (csrc/gemm.cu) and a counter-based inter-layer dropout mask; everything else (float32, hidden sizes
that are not a multiple of 32 or above 512, an initial state, time-major input) is handed to torch's library LSTM, exactly as
Decoder.run_rnn does ("Precision": the float32 contractions are a library validation path).
"""
from __future__ import annotations

import torch
from torch import nn

from .. import _lib as L
from .. import lstm as plstm

_instances = 0


class LSTM(nn.Module):
    def __init__(self, input_size, hidden_size, num_layers=1, bias=True, batch_first=False, dropout=0.0, bidirectional=False,
                 seed: int = None):
        super().__init__()
        global _instances
        _instances += 1
        self.input_size, self.hidden_size, self.num_layers = int(input_size), int(hidden_size), int(num_layers)
        self.bias, self.batch_first, self.dropout, self.bidirectional = bool(bias), bool(batch_first), float(dropout), bool(bidirectional)
        ref = nn.LSTM(self.input_size, self.hidden_size, self.num_layers, bias=self.bias, batch_first=self.batch_first,
                      dropout=self.dropout, bidirectional=self.bidirectional)           # torch's own initialisation and names
        self._names = [n for n, _ in ref.named_parameters()]
        for n, p in ref.named_parameters():
            self.register_parameter(n, nn.Parameter(p.detach().clone()))
        self.dropout_seed = 0x4C53544D + _instances if seed is None else int(seed)
        self.dropout_calls = 0

    def _layer_params(self, layer, sfx=""):
        return [getattr(self, f"{k}_l{layer}{sfx}") for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]

    def _persistent_ok(self, x, hx):
        if hx is not None or not self.batch_first or not self.bias or x.dim() != 3:
            return False
        return plstm.supported(x, self.hidden_size) if self.bidirectional else plstm.supported_uni(x, self.hidden_size)

    def forward(self, x, hx=None):
        L.require_cuda(x)                                    # like every module of the package: CUDA tensors only, no CPU path
        if not self._persistent_ok(x, hx):
            flat = [getattr(self, n).to(x.dtype) for n in self._names]
            nd = 2 if self.bidirectional else 1
            if hx is None:
                bdim = x.shape[0] if self.batch_first else x.shape[1]
                z = x.new_zeros(nd * self.num_layers, bdim, self.hidden_size)
                hx = (z, z)
            out, h, c = torch._VF.lstm(x, hx, flat, self.bias, self.num_layers, self.dropout, self.training, self.bidirectional, self.batch_first)
            return out, (h, c)
        need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        H = self.hidden_size
        h_n = []
        for layer in range(self.num_layers):
            drop = None
            if self.dropout > 0 and self.training and layer > 0:
                drop = (self.dropout, self.dropout_seed, self.dropout_calls * self.num_layers + layer - 1, None)
            if self.bidirectional:
                x = plstm.bilstm_layer(x, *self._layer_params(layer), *self._layer_params(layer, "_reverse"), training=need_grad,
                                       input_dropout=drop)
                h_n += [x[:, -1, :H], x[:, 0, H:]]
            else:
                x = plstm.lstm_layer(x, *self._layer_params(layer), training=need_grad, input_dropout=drop)
                h_n.append(x[:, -1, :])
        if self.dropout > 0 and self.training:
            self.dropout_calls += 1
        # c_n is not kept by the inference kernel; the reference's call sites read only the output (MD_VAE/model.py:116 `[0]`)
        return x, (torch.stack(h_n, 0), None)
