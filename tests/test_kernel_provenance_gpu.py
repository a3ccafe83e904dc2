"""GPU: WHICH kernels a training step launches (torch.profiler kernel names).

bf16 (the product path, BASELINE configs[1]): every contraction -- LSTM recurrences, the time-parallel GEMMs, the dense
layers -- the loss / dropout / normaliser / front-end kernels and the optimiser are kernels of libmlvae_b200.so; NO cuBLAS
(nvjet / cutlass / gemm / gemv), NO cuDNN (RNN_ / cudnn) kernel may appear.
float32 is the reference's own precision (md_model.py:77-86) and is kept as a VALIDATION path: its contractions are library
kernels (cuDNN LSTM, cuBLAS SIMT sgemm -- tensor cores would break the 1e-5 parity), so the float32 golden tests exercise the
repo's loss / mask / normaliser / front-end kernels and the host glue, not a repo GEMM.  This test pins that statement."""
import pytest
import torch

pytestmark = pytest.mark.gpu

LIBRARY_MARKERS = ("nvjet", "cutlass", "cublas", "sgemm", "gemv", "xmma", "cudnn", "RNN_", "ampere_", "sm90_", "sm100_")


def _kernel_names(cuda, dtype):
    from torch.profiler import ProfilerActivity, profile
    from ml_vae_b200.features import Fbank
    from ml_vae_b200.modules import Decoder, VanillaVAE
    from ml_vae_b200.normalizer import InputNormalization
    from ml_vae_b200.train_step import TrainStep
    torch.manual_seed(1)
    B, n = 16, 16000
    enc = VanillaVAE([80, 64, 64], 64).to(cuda)
    dec = Decoder(64, 128, 2, 0.15, [256, 64, 64, 80]).to(cuda)
    ts = TrainStep(Fbank(deltas=False, hop_length=10, n_mels=80), InputNormalization().to(cuda), enc, dec,
                   {"kld_weight": 0.001, "batch_size": B}, compute_dtype=dtype)
    wav = 0.1 * torch.randn(B, n, device=cuda)
    lens = torch.full((B,), n, dtype=torch.int32, device=cuda)
    ts.step(wav, lens)                                   # warm-up (lazy initialisation kernels)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ts.step(wav, lens)
        torch.cuda.synchronize()
    return [e.key for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA or getattr(e, "self_device_time_total", 0) > 0]


def test_bf16_step_has_no_library_contraction(cuda):
    names = _kernel_names(cuda, torch.bfloat16)
    lib = [n for n in names if any(m in n for m in LIBRARY_MARKERS)]
    assert not lib, lib
    ours = " ".join(n for n in names if "mlvae" in n)
    for must in ("lstm_fwd_kernel", "lstm_bwd_kernel", "gemm_bf16_kernel", "linear_fwd_kernel", "dense_bwd_prep_kernel", "adam_step_kernel",
                 "reparam_kl_fwd_kernel", "recon_fwd_kernel", "logmel", "dropout_kernel"):
        assert must in ours, must


def test_fp32_step_is_a_library_validation_path(cuda):
    names = _kernel_names(cuda, torch.float32)
    joined = " ".join(names)
    assert any(m in joined for m in ("RNN_", "cudnn")), "float32 LSTM is expected on cuDNN"
    assert any(m in joined for m in ("sgemm", "gemm", "nvjet", "cutlass")), "float32 Linear is expected on cuBLAS"
    assert "lstm_fwd_kernel" not in joined and "gemm_bf16_kernel" not in joined
    for must in ("reparam_kl_fwd_kernel", "recon_fwd_kernel", "logmel", "adam_step_kernel"):      # the repo kernels the fp32 tests DO cover
        assert must in joined, must


def test_lstm_module_and_decoder_run_on_repo_kernels(cuda):
    """The torch.nn.LSTM drop-in (forward-only, bf16) launches the one-direction recurrence instantiation and the TMA GEMM, no cuDNN /
    cuBLAS kernel; the label decoder is one md_decode_kernel launch."""
    from torch.profiler import ProfilerActivity, profile
    from ml_vae_b200.modules import LSTM
    from ml_vae_b200.utils import decode_utils as du
    m = LSTM(128, 256, 2, batch_first=True, dropout=0.15).to(cuda).train()
    x = torch.randn(8, 40, 128, device=cuda).bfloat16().requires_grad_(True)
    m(x)[0].float().sum().backward()
    logs = [torch.log(torch.rand(s, device=cuda).clamp_min(1e-5)) for s in ((4, 30, 9, 2), (4, 30, 2), (4, 30, 2), (9, 2))]
    y = torch.randint(0, 9, (4, 6), device=cuda)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        m(x)[0].float().sum().backward()
        du.decode_from_logs(*logs, y, torch.full((4,), 30), torch.full((4,), 6), device=cuda)
        torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages()]
    lib = [n for n in names if any(k in n for k in LIBRARY_MARKERS)]
    assert not lib, lib
    joined = " ".join(names)
    for must in ("lstm_fwd_kernel", "lstm_bwd_kernel", "gemm_bf16_kernel", "dropout_kernel", "md_decode_kernel"):
        assert must in joined, must
