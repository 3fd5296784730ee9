"""Backward building blocks of the head against plain fp32 / autograd references."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from cmpc_refseg_b200 import _lib as L
    dev = torch.device("cuda:0")
    return L, L.lib(), dev, torch.cuda.current_stream(dev).cuda_stream


@pytest.mark.parametrize("M,I,J,splits", [(64, 128, 256, 1), (200, 72, 40, 0), (1000, 264, 520, 3), (25600, 1008, 1000, 0),
                                          (51200, 504, 2048, 0), (7, 8, 8, 0)])
def test_gemm_atb_weight_gradient(env, M, I, J, splits):
    """dW[cin, cout] = X^T dY (1x1 conv wgrad in the TF layout), accumulated on top of what is already in the buffer."""
    L, lib, dev, st = env
    lda, ldb, ldo = (I + 63) // 64 * 64, (J + 63) // 64 * 64 + 8, J + 4
    a = torch.zeros(M, lda, device=dev, dtype=torch.float16); a[:, :I] = (torch.randn(M, I, device=dev) * 0.5).half()
    b = torch.zeros(M, ldb, device=dev, dtype=torch.float16); b[:, :J] = (torch.randn(M, J, device=dev) * 0.5).half()
    a[:, I:] = 9.0; b[:, J:] = 9.0                                   # pad columns must not leak into the result
    out = torch.full((I, ldo), 0.25, device=dev)
    L.check(lib.cmpc_gemm_atb_f16(a.data_ptr(), lda, I, b.data_ptr(), ldb, J, M, out.data_ptr(), ldo, splits, st), "atb")
    torch.cuda.synchronize()
    ref = a[:, :I].double().t() @ b[:, :J].double() + 0.25
    err = (out[:, :J].double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f"atb M={M} I={I} J={J}: max-abs {err:.3e} (ref max {scale:.3e})")
    assert err <= 2e-5 * scale * max(1.0, (M / 1000) ** 0.5) + 1e-4
    assert (out[:, J:] == 0.25).all()
