"""Drop-in for ``speechbrain.processing.features.InputNormalization(norm_type='global')`` as declared at
models/test_vanilla_vae/model.yaml:14-15 and called at models/test_vanilla_vae/model.py:24-25
[arithmetic: SB-recall, SpeechBrain 0.5.x]:

  per utterance: mean and unbiased std over the round(len * T) valid frames (std floored at 1e-10), averaged over
  the batch; running global statistics updated with weight 1 / (count + 1) while epoch < update_until_epoch (3);
  output (x - glob_mean) / glob_std.

SpeechBrain loops over the batch in python with an .int() sync per utterance; here three small kernels
(csrc/norm.cu) do it with the running state (count, mean, std) resident on the device: no host sync, and the call is
capturable in a CUDA graph.  SURVEY.md section 8f-2.
"""
from __future__ import annotations

import torch

from . import _lib as L


class InputNormalization(torch.nn.Module):
    def __init__(self, mean_norm=True, std_norm=True, norm_type="global", avg_factor=None,
                 requires_grad=False, update_until_epoch=3, sync_stats: bool = False, process_group=None):
        super().__init__()
        if norm_type != "global" or not (mean_norm and std_norm) or avg_factor is not None or requires_grad:
            raise NotImplementedError("only InputNormalization(norm_type='global') with SpeechBrain defaults "
                                      "(the reference's configuration, model.yaml:14-15) is implemented")
        self.update_until_epoch = update_until_epoch
        # data parallel (optional): all-reduce the batch statistics (2 D floats) so that every rank keeps the
        # running statistics of the GLOBAL batch; off by default = per-process statistics, what the reference does under DDP
        self.sync_stats = bool(sync_stats)
        self.process_group = process_group
        self._avg = None
        self._state = None          # {count, pad[3], glob_mean[D], glob_std[D]} float32 on the device
        self._scratch = None
        self._dim = None

    # running statistics, as views of the device state
    @property
    def count(self) -> int:
        return 0 if self._state is None else int(self._state[0].item())

    @property
    def glob_mean(self):
        return None if self._state is None else self._state[4:4 + self._dim]

    @property
    def glob_std(self):
        return None if self._state is None else self._state[4 + self._dim:4 + 2 * self._dim]

    # checkpointing: the running statistics are lazily sized device state, carried as the module's extra state
    def get_extra_state(self):
        return {"dim": self._dim, "state": None if self._state is None else self._state.detach().cpu().clone()}

    def set_extra_state(self, st):
        self._dim = st["dim"]
        dev = self._state.device if self._state is not None else (torch.device("cuda", torch.cuda.current_device())
                                                                   if torch.cuda.is_available() else torch.device("cpu"))
        self._state = None if st["state"] is None else st["state"].to(dev).clone()

    @torch.no_grad()
    def forward(self, x, lengths, epoch=0, out_dtype=None):
        L.require_cuda(x)
        if x.dim() != 3:
            raise ValueError(f"expected (batch, time, features), got {tuple(x.shape)}")
        B, T, D = x.shape
        xf = x.float().contiguous()
        lens = lengths.to(device=x.device, dtype=torch.float32).contiguous()
        if self._state is None or self._dim != D or self._state.device != x.device:
            self._state = torch.zeros(L.lib().mlvae_norm_state_bytes(D) // 4, dtype=torch.float32, device=x.device)
            self._dim = D
        need = L.lib().mlvae_norm_scratch_bytes(B, D) // 4
        if self._scratch is None or self._scratch.numel() < need or self._scratch.device != x.device:
            self._scratch = torch.empty(need, dtype=torch.float32, device=x.device)
        out = torch.empty(B, T, D, dtype=out_dtype or x.dtype, device=x.device)
        if not self.training and self.count == 0:
            raise RuntimeError("InputNormalization(global) used in eval mode before any statistics were seen")
        world = 1
        if self.sync_stats and self.training and torch.distributed.is_available() and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size(self.process_group)
        if world > 1:
            if self._avg is None or self._avg.numel() != 2 * D or self._avg.device != x.device:
                self._avg = torch.empty(2 * D, dtype=torch.float32, device=x.device)
            L.check(L.lib().mlvae_global_norm_batch_avg(L.ptr(xf), L.ptr(lens), B, T, D, L.ptr(self._scratch), L.ptr(self._avg), L.stream_ptr()),
                    "mlvae_global_norm_batch_avg", kernels=2)
            torch.distributed.all_reduce(self._avg, op=torch.distributed.ReduceOp.SUM, group=self.process_group)
            L.check(L.lib().mlvae_global_norm_from_avg(L.ptr(xf), B, T, D, L.ptr(self._avg), 1.0 / world, int(epoch < self.update_until_epoch),
                                                       L.ptr(self._state), L.ptr(out), L.dtype_code(out), L.stream_ptr()),
                    "mlvae_global_norm_from_avg", kernels=2)
            return out
        L.check(L.lib().mlvae_global_norm(L.ptr(xf), L.ptr(lens), B, T, D, int(self.training),
                                          int(epoch < self.update_until_epoch), L.ptr(self._state), L.ptr(self._scratch),
                                          L.ptr(out), L.dtype_code(out), L.stream_ptr()), "mlvae_global_norm",
                kernels=3 if self.training else 1)
        return out
