"""Drop-in for the reference's modules/decoder.py:10-53 (Decoder).

Same constructor kwargs (input_size, rnn_hidden_size, rnn_num_layers, rnn_dropout,
fc_sizes, loss_type='likelihood'), checkpoint keys (rnn.weight_ih_l0[_reverse]...,
mean_fc.blocks.{0,2,4}.*, log_var_fc.blocks.{0,2,4}.*) and forward() dict
{'mean', 'log_var', 'losses': {'recon_loss': unreduced (B, T, D)}}; an unknown
loss_type raises ValueError exactly like decoder.py:51.

The reconstruction loss (and, with ``lens=``, its length-masked mean) is one fused
kernel; the dead Normal.log_prob of decoder.py:45-47 is not computed.  bf16 activations run the
biLSTM on the persistent tcgen05 recurrence (lstm.py); float32 stays on cuDNN.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import lstm, mlp_chain, ops
from ..dense import direct_chain, linear, linear_chain, linear_direct
from ._params import attach, torch_default_linear, torch_default_lstm


class Decoder(nn.Module):
    def __init__(self, input_size, rnn_hidden_size, rnn_num_layers, rnn_dropout, fc_sizes, loss_type="likelihood",
                 materialize_loss: bool = True):
        super().__init__()
        self.input_size = int(input_size)
        self.hidden = int(rnn_hidden_size)
        self.num_layers = int(rnn_num_layers)
        self.rnn_dropout = float(rnn_dropout)
        self.fc_sizes = [int(s) for s in fc_sizes]
        self.loss_type = loss_type
        self.materialize_loss = materialize_loss
        self.use_persistent_lstm = True
        # inter-layer dropout mask: counter-based Philox stream (csrc/dropout.cu) keyed by (seed, call count, layer)
        # [+ a device step counter when the training step is replayed as a CUDA graph]
        self.dropout_seed = 123456 ^ 0x5DEECE66D
        self.dropout_calls = 0
        self.dropout_offset_dev = None
        self.direct_param_grads = False       # train_step.py: LSTM gradients accumulate straight into the flat bucket
        self.defer_weight_grads = False       # train_step.py: upper layers' dW GEMMs run beside the recurrence of the layer below
        self.top_layer_grad_hook = None      # callable(): runs in backward once the top LSTM layer's and the heads' gradients exist
        self._rnn_names = []
        for name, p in torch_default_lstm(self.input_size, self.hidden, self.num_layers):
            attach(self, f"rnn.{name}", p)
            self._rnn_names.append(name)
        for head in ("mean_fc", "log_var_fc"):
            for i in range(len(self.fc_sizes) - 1):
                w, b = torch_default_linear(self.fc_sizes[i], self.fc_sizes[i + 1])
                attach(self, f"{head}.blocks.{2 * i}.weight", w)
                attach(self, f"{head}.blocks.{2 * i}.bias", b)

    # -- flat-arena binding (train_step.FlatArena) -------------------------------------------------------------
    def adjacent_param_groups(self):
        """Parameters the arena should lay out back to back: the first layers of the two heads run as ONE stacked GEMM."""
        (wm, bm), (wv, bv) = self._head("mean_fc"), self._head("log_var_fc")
        if len(wm) > 1 and wm[0].shape == wv[0].shape:
            return [[wm[0], wv[0]], [bm[0], bv[0]]]
        return []

    def bind_arena(self, arena):
        """Use the arena's bf16 shadow weights directly and accumulate the dense gradients in its bucket (bf16 path only)."""
        (wm, bm), (wv, bv) = self._head("mean_fc"), self._head("log_var_fc")
        self._direct = None
        if len(wm) > 1 and wm[0].shape == wv[0].shape:
            head0 = arena.linear_views([wm[0], wv[0]], [bm[0], bv[0]])
            mean = [arena.linear_views([w], [b]) for w, b in zip(wm[1:], bm[1:])]
            logv = [arena.linear_views([w], [b]) for w, b in zip(wv[1:], bv[1:])]
            if head0 is not None and all(v is not None for v in mean + logv):
                self._direct = (head0, mean, logv)

    def _head(self, name):
        blocks = getattr(self, name).blocks._modules
        n = len(self.fc_sizes) - 1
        return [blocks[str(2 * i)].weight for i in range(n)], [blocks[str(2 * i)].bias for i in range(n)]

    def run_rnn(self, x):
        """decoder.py:14-15,22: 2-layer bidirectional LSTM, batch_first, inter-layer dropout.
        bf16 activations run on the persistent tcgen05 recurrence (csrc/lstm.cu); float32 (and hidden
        sizes that are not a multiple of 32) go through cuDNN."""
        if self.use_persistent_lstm and lstm.supported(x, self.hidden):
            need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.rnn.parameters()))
            for layer in range(self.num_layers):
                ps = [getattr(self.rnn, f"{kind}_l{layer}{sfx}") for sfx in ("", "_reverse")
                      for kind in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
                # train_step.py (data parallel): once the backward recurrence of the layer BELOW the top one is enqueued, the
                # gradients of the top layer and of the heads are final -> their all-reduce overlaps this layer's GEMMs
                hook = self.top_layer_grad_hook if (layer + 2 == self.num_layers and need_grad) else None
                # nn.LSTM's inter-layer dropout (decoder.py:14-15) acts on the INPUT of every layer but the first: handed to the
                # layer, which applies the counter-based mask going in and folds it into its input-gradient GEMM coming back
                drop = None
                if self.rnn_dropout > 0 and self.training and layer > 0:
                    drop = (self.rnn_dropout, self.dropout_seed, self._dropout_offset(layer - 1), self.dropout_offset_dev)
                x = lstm.bilstm_layer(x, *ps, training=need_grad, direct_grads=self.direct_param_grads, after_recurrence=hook,
                                      input_dropout=drop, defer_weight_grads=self.defer_weight_grads and layer > 0)
            if self.rnn_dropout > 0 and self.training:
                self.dropout_calls += 1
            return x
        z = x.new_zeros(2, x.shape[0], self.hidden)
        if self.rnn_dropout > 0 and self.training and self.num_layers > 1:
            # library LSTM one layer at a time with OUR counter-based dropout between the layers (cuDNN's own dropout draws
            # from a stateful generator nothing else can reproduce)
            for layer in range(self.num_layers):
                flat = [getattr(self.rnn, f"{kind}_l{layer}{sfx}").to(x.dtype) for sfx in ("", "_reverse")
                        for kind in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
                x, _, _ = torch._VF.lstm(x, (z, z), flat, True, 1, 0.0, self.training, True, True)
                if layer + 1 < self.num_layers:
                    x = ops.dropout(x, self.rnn_dropout, self.dropout_seed, self._dropout_offset(layer), self.dropout_offset_dev)
            self.dropout_calls += 1
            return x
        flat = [getattr(self.rnn, n).to(x.dtype) for n in self._rnn_names]
        z = x.new_zeros(2 * self.num_layers, x.shape[0], self.hidden)
        out, _, _ = torch._VF.lstm(x, (z, z), flat, True, self.num_layers, 0.0, self.training, True, True)
        return out

    def _dropout_offset(self, layer: int) -> int:
        # low word = device step counter (added by the kernel when present), high word = host call count and layer
        k = self.dropout_calls * self.num_layers + layer
        return (k << 32) if self.dropout_offset_dev is not None else k

    def forward(self, sampled_h, target_feats, lens=None):
        if self.loss_type not in ("likelihood", "mse"):
            raise ValueError(f"Invalid loss type: {self.loss_type}")
        rnn_out = self.run_rnn(sampled_h)
        (wm, bm), (wv, bv) = self._head("mean_fc"), self._head("log_var_fc")
        direct = getattr(self, "_direct", None)
        if direct is not None and rnn_out.dtype == torch.bfloat16:
            head0, mviews, vviews = direct
            n0 = wm[0].shape[0]
            h0 = linear_direct(rnn_out, head0, leaky=True)
            if (len(mviews) == 2 and len(vviews) == 2 and mviews[0][0].shape == vviews[0][0].shape and mviews[1][0].shape == vviews[1][0].shape
                    and mlp_chain.supported(n0, mviews[0][0].shape[0], mviews[1][0].shape[0])):
                # the tails of BOTH heads (64 -> 64 -> D each) in one fused launch per pass (csrc/mlp_chain.cu)
                mean, log_var = mlp_chain.chain2(h0, [mviews[0], vviews[0]], [mviews[1], vviews[1]], act_b=False)
            else:
                mean = direct_chain(h0[..., :n0], mviews)
                log_var = direct_chain(h0[..., n0:], vviews)
        elif len(wm) > 1 and wm[0].shape == wv[0].shape:
            # both heads read the (B, T, 2H) LSTM output: their first layers run as ONE GEMM against the stacked
            # weight, so the widest activation of the model is read once (decoder.py:24-25 reads it twice)
            n0 = wm[0].shape[0]
            h0 = linear(rnn_out, torch.cat([wm[0], wv[0]], 0), torch.cat([bm[0], bv[0]], 0), leaky=True)
            mean = linear_chain(h0[..., :n0], wm[1:], bm[1:])
            log_var = linear_chain(h0[..., n0:], wv[1:], bv[1:])
        else:
            mean = linear_chain(rnn_out, wm, bm)
            log_var = linear_chain(rnn_out, wv, bv)
        elem, red = ops.recon_loss(mean, log_var, target_feats, lens=lens, loss_type=self.loss_type,
                                   want_elem=self.materialize_loss, want_mean=lens is not None)
        out = {"mean": mean, "log_var": log_var, "losses": {"recon_loss": elem}}
        if lens is not None:
            out["recon_loss"] = red
        return out

    def compute_recon_loss(self, mean, log_var, target):
        return ops.recon_loss(mean, log_var, target, loss_type=self.loss_type, want_elem=True, want_mean=False)[0]
