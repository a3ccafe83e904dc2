"""Stand-alone probe of the persistent LSTM forward (run under `timeout`)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ml_vae_b200 import _lib as L

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def run(B, T, In, H, time_it=False):
    torch.manual_seed(B * 7 + T + H)
    lstm = torch.nn.LSTM(In, H, 1, bidirectional=True, batch_first=True).to(dev)
    with torch.no_grad():
        for p_ in lstm.parameters():
            p_.copy_(p_.bfloat16().float())
    x = torch.randn(B, T, In, device=dev).bfloat16().float()
    with torch.no_grad():
        ref, (hn, cn) = lstm(x)
    wih = torch.cat([lstm.weight_ih_l0, lstm.weight_ih_l0_reverse], 0)                   # (8H, In)
    bias = torch.cat([lstm.bias_ih_l0 + lstm.bias_hh_l0, lstm.bias_ih_l0_reverse + lstm.bias_hh_l0_reverse], 0)
    from ml_vae_b200.lstm import _gate_perm
    perm, _ = _gate_perm(H, dev)
    P = (x @ wih[perm].t() + bias[perm]).bfloat16().contiguous()                             # (B, T, 8H) = (B,T,2,H,4)
    whh = torch.stack([lstm.weight_hh_l0, lstm.weight_hh_l0_reverse], 0).bfloat16().contiguous()
    Y = torch.full((B, T, 2 * H), float("nan"), device=dev, dtype=torch.bfloat16)
    C = torch.empty(B, T, 2 * H, device=dev)
    scratch = torch.empty(L.lib().mlvae_lstm_scratch_bytes(B, H), dtype=torch.uint8, device=dev)
    P0 = P.clone()
    L.check(L.lib().mlvae_lstm_fwd(L.ptr(P), L.ptr(whh), L.ptr(Y), L.ptr(C), B, T, H, 1, L.ptr(scratch), L.stream_ptr()), "lstm_fwd")
    torch.cuda.synchronize()
    err = float((Y.float() - ref).abs().max() / ref.abs().max())
    cerr = float((C[:, -1, :H] - cn[0]).abs().max() / cn[0].abs().max())
    print(f"B={B} T={T} In={In} H={H}: y rel err {err:.3e}  c_T rel err {cerr:.3e}", "OK" if err < 2e-2 else "MISMATCH", flush=True)
    if time_it:
        prof = torch.zeros(128, dtype=torch.int64, device=dev)
        L.check(L.lib().mlvae_debug_set_profile_buffer(L.ptr(prof)), "prof")
        L.check(L.lib().mlvae_lstm_fwd(L.ptr(P), L.ptr(whh), L.ptr(Y), L.ptr(C), B, T, H, 1, L.ptr(scratch), L.stream_ptr()), "lstm_fwd")
        torch.cuda.synchronize()
        L.check(L.lib().mlvae_debug_set_profile_buffer(None), "prof")
        names = ["exchange wait", "B tile + hand-off", "mma completion", "ld + gates + publish + stores",
                 "issuer 0: wait", "issuer 0: issue + commit", "issuer 1: wait", "issuer 1: issue + commit"]
        pc = prof.cpu().tolist()
        print("   cycles/step:", {n: round(v / T) for n, v in zip(names, pc)}, "total", round(sum(pc[:4]) / T))
        print("   per CTA of group 0 (gate warp 0, chain 0) [wait, B tile, mma, gates]:")
        for x in range(H // 32):
            print("     cta", x, [round(v / T) for v in pc[12 + 4 * x: 16 + 4 * x]])
        print("   per gate warp of CTA 0 chain 0:")
        for w in range(1, 8):
            print("     warp", w, [round(v / T) for v in pc[76 + 4 * w: 80 + 4 * w]])
        for sv in (0, 1):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                L.check(L.lib().mlvae_lstm_fwd(L.ptr(P), L.ptr(whh), L.ptr(Y), L.ptr(C) if sv else None, B, T, H, sv, L.ptr(scratch), L.stream_ptr()), "lstm_fwd")
            b.record(); torch.cuda.synchronize()
            print(f"   save={sv}: {a.elapsed_time(b) / 3:.3f} ms")
        for _ in range(2):
            P.copy_(P0)
            L.check(L.lib().mlvae_lstm_fwd(L.ptr(P), L.ptr(whh), L.ptr(Y), L.ptr(C), B, T, H, 1, L.ptr(scratch), L.stream_ptr()), "lstm_fwd")
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            L.check(L.lib().mlvae_lstm_fwd(L.ptr(P), L.ptr(whh), L.ptr(Y), L.ptr(C), B, T, H, 1, L.ptr(scratch), L.stream_ptr()), "lstm_fwd")
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"   persistent kernel {ms:.3f} ms ({ms / T * 1e3:.2f} us/step)", flush=True)
    return err < 2e-2


import sys
if len(sys.argv) > 1:                      # python tests/lstm_probe.py B T  -> instrumented timing of that shape only
    run(int(sys.argv[1]), int(sys.argv[2]), 64, 512, time_it=True)
    sys.exit(0)
ok = True
ok &= run(4, 6, 16, 32)
ok &= run(16, 20, 24, 64)
ok &= run(20, 33, 64, 128)
ok &= run(64, 50, 64, 512)
ok &= run(64, 500, 64, 512, time_it=True)
ok &= run(16, 2000, 64, 512, time_it=True)
print("ALL OK" if ok else "FAILED")
