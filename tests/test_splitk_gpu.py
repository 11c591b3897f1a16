"""Split-K tail scheduling of the tcgen05 GEMM: a ragged last wave (456 tiles on 148 SMs) is split along K over idle
SMs through an fp32 workspace; results must equal the unsplit kernel's and the tile counters must come back zeroed."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

import algonauts2025_b200  # noqa: E402,F401
from algonauts2025_b200 import ops  # noqa: E402


def _ref_close(out, ref):
    err = (out.float() - ref).abs()
    assert bool((err <= 1e-2 * float(ref.abs().max()) + 1e-2 * ref.abs()).all()), float(err.max())


@pytest.mark.parametrize("shape", [(4768, 3072, 3072), (4768, 3072, 12288), (1100, 2904, 1000)])
@pytest.mark.parametrize("b_mn", [False, True])
def test_splitk_matches_unsplit_and_restores_workspace(shape, b_mn):
    m, n, k = shape
    torch.manual_seed(0)
    A = (torch.randn(m, k, device="cuda") / math.sqrt(k)).bfloat16()
    W = torch.randn(n, k, device="cuda").bfloat16()
    b_op = ops.mnmajor(W.t().contiguous()) if b_mn else ops.kmajor(W)
    bias, res, rs = torch.randn(n, device="cuda"), torch.randn(m, n, device="cuda"), torch.rand(n, device="cuda")
    ref = A.float() @ W.float().t() + bias + res * rs
    outs = []
    for split in (True, False, True):
        ops.SPLITK = split
        out = torch.full((m, n), float("nan"), device="cuda")
        ops.gemm(ops.kmajor(A), b_op, out, m, n, k, ldd=n, bias=bias, epilogue=ops.EPI_RESIDUAL, res=res, ld_res=n, rscale=rs)
        outs.append(out)
    ops.SPLITK = True
    torch.cuda.synchronize()
    for o in outs:
        _ref_close(o, ref)
    # same inputs, different reduction order in the tail tiles only: tiny fp32 differences allowed
    assert float((outs[0] - outs[1]).abs().max()) <= 1e-3 * float(ref.abs().max())
    ws = ops._splitk_workspace(torch.device("cuda", torch.cuda.current_device()))
    assert int(ws[:2560].count_nonzero()) == 0  # arrival / departure counters are reset by the last slice
    # bf16 output through the split path as well
    outb = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    ops.gemm(ops.kmajor(A), b_op, outb, m, n, k, ldd=n)
    _ref_close(outb, A.float() @ W.float().t())
