"""Dense MLP head on the tcgen05 tensor cores (hpsx_mlp_*, include/hpsx.h) against a plain PyTorch
fp32 reference of the same op.  The kernel computes with bf16 operands and fp32 accumulation, so the reference uses
the bf16-rounded operands in fp32 and the tolerance is the bf16 rounding of the hidden activations (2^-8 relative,
written below).  Reference model: samples/hps-triton-ensemble/01_model_training.ipynb cells 7,11
(fc_1 256 -> fc_2 128 -> fc_3 1 on [batch, slot_num * embed_vec_size], no activations)."""
import numpy as np
import pytest

import hugectr_backend_b200 as hb
from oracle import hps_oracle as O

pytestmark = pytest.mark.gpu

RTOL, ATOL = 2e-2, 2e-2  # bf16 operands and hidden activations: 8 bits of mantissa


def _torch():
    import torch

    return torch


def reference(torch, x, weights, biases, relu):
    """fp32 math on bf16-rounded operands; hidden activations rounded to bf16 like the kernel stores them."""
    h = x.to(torch.bfloat16).to(torch.float32)
    L = len(weights)
    for l, w in enumerate(weights):
        wq = torch.from_numpy(w).cuda().to(torch.bfloat16).to(torch.float32)
        h = h @ wq.t()
        if biases[l] is not None:
            h = h + torch.from_numpy(biases[l]).cuda()
        if relu[l]:
            h = torch.relu(h)
        if l + 1 < L:
            h = h.to(torch.bfloat16).to(torch.float32)
    return h


@pytest.mark.parametrize("batch,dims,relu", [
    (128, [64, 128], [0]),                       # one tile, one k-block
    (128, [256, 128], [0]),                      # ring wrap-around (4 k-blocks, 3 stages)
    (1000, [48, 256, 128, 1], [0, 0, 0]),        # the sample's dense model (3 slots x 16), ragged batch
    (4096, [3328, 1024, 512, 256, 1], [1, 1, 1, 0]),  # Criteo-shape head: 26 slots x 128
    (333, [136, 200, 72], [1, 0]),               # N and K tails (not multiples of the tile)
])
def test_mlp_matches_torch_reference(cuda_device, batch, dims, relu):
    torch = _torch()
    torch.backends.cuda.matmul.allow_tf32 = False
    rng = np.random.default_rng(batch + len(dims))
    weights = [(rng.standard_normal((dims[l + 1], dims[l])) / np.sqrt(dims[l])).astype(np.float32) for l in range(len(dims) - 1)]
    biases = [rng.standard_normal(dims[l + 1]).astype(np.float32) * 0.1 if l % 2 == 0 else None for l in range(len(dims) - 1)]
    mlp = hb.DenseMlp(0, weights, biases, relu)
    x = torch.from_numpy(rng.standard_normal((batch, dims[0])).astype(np.float32)).cuda()
    out = torch.full((batch, dims[-1]), float("nan"), device="cuda")
    mlp.forward(x, batch, out)
    torch.cuda.synchronize()
    want = reference(torch, x, weights, biases, relu)
    assert torch.isfinite(out).all()
    torch.testing.assert_close(out, want, rtol=RTOL, atol=ATOL)
    mlp.close()


def test_lookup_feeds_the_dense_head_in_place(cuda_device):
    """The ensemble of the sample in one process: HPS lookup ([batch, 3] keys -> [batch, 48] vectors in device memory)
    followed by the dense model reading that buffer directly (no LOOKUP_VECTORS hand-off through Triton)."""
    torch = _torch()
    slots, dim, batch, rows = 3, 16, 2048, 30_000
    hps = hb.HPS(num_partitions=4)
    hps.add_model(hb.ModelParams("naive_dnn", batch, [dim], [slots], [0.0], hit_rate_threshold=1.0, cache_size_percentage=0.5))
    hps.load_table_procedural("naive_dnn", 0, rows, 7)
    hps.create_embedding_cache("naive_dnn")
    table = O.NumpyTable(dim, 0.0)
    table.fill_procedural(rows, 7)
    rng = np.random.default_rng(0)
    weights = [np.ones((256, 48), np.float32) * 0.02, np.ones((128, 256), np.float32) * 0.01, np.ones((1, 128), np.float32) * 0.05]
    mlp = hb.DenseMlp(0, weights, None, None)
    keys = rng.integers(0, rows, size=batch * slots)
    vec = torch.empty((batch * slots, dim), device="cuda")
    s = hps.session("naive_dnn", 0)
    s.lookup([keys], [vec], [len(keys)])
    logit = torch.empty((batch, 1), device="cuda")
    mlp.forward(vec, batch, logit, stream=s.stream)  # same stream as the lookup: no host hand-off
    torch.cuda.synchronize()
    x = torch.from_numpy(table.lookup(keys).reshape(batch, slots * dim)).cuda()
    want = reference(torch, x, weights, [None] * 3, [0, 0, 0])
    torch.testing.assert_close(logit, want, rtol=RTOL, atol=ATOL)


def test_rejects_bad_shapes(cuda_device):
    with pytest.raises(hb.HpsxError, match="multiple of 8"):
        hb.DenseMlp(0, [np.ones((16, 12), np.float32)])  # input width not a multiple of 8
    with pytest.raises(hb.HpsxError, match="multiple of 8"):
        hb.DenseMlp(0, [np.ones((1, 16), np.float32), np.ones((8, 1), np.float32)])  # a single unit cannot feed a GEMM layer
    with pytest.raises(ValueError):
        hb.DenseMlp(0, [np.ones((8, 16), np.float32), np.ones((8, 16), np.float32)])  # widths do not chain


@pytest.mark.parametrize("pagelock", [False, True])
def test_bf16_mirror_is_the_rounded_fp32_output_and_feeds_the_head(cuda_device, pagelock):
    """The lookup writes a bf16 copy of its rows from the same kernels (hits: probe+gather, misses: pull / merge):
    bit-exact round-to-nearest-even of the fp32 rows, in both miss paths, and the dense head gives the same logits
    from it as from the fp32 output (no conversion pass)."""
    torch = _torch()
    slots, dim, batch, rows = 26, 128, 1500, 60_000
    n = batch * slots
    hps = hb.HPS(num_partitions=4)
    hps.add_model(hb.ModelParams("m", batch, [dim], [slots], [0.75], hit_rate_threshold=0.5, cache_size_percentage=0.3,
                                 enable_pagelock=pagelock))
    hps.load_table_procedural("m", 0, rows, 3)
    hps.create_embedding_cache("m")
    table = O.NumpyTable(dim, 0.75)
    table.fill_procedural(rows, 3)
    s = hps.session("m", 0)
    rng = np.random.default_rng(8)
    dims = [slots * dim, 512, 1]
    weights = [(rng.standard_normal((dims[l + 1], dims[l])) / np.sqrt(dims[l])).astype(np.float32) for l in range(2)]
    mlp = hb.DenseMlp(0, weights, None, [1, 0])
    for it in range(3):
        keys = rng.integers(-3, rows + 3, size=n)
        out = torch.full((n, dim), float("nan"), device="cuda")
        mirror = torch.zeros((n, dim), dtype=torch.bfloat16, device="cuda")
        if it == 1:
            dk = torch.from_numpy(keys).cuda()
            torch.cuda.synchronize()
            s.lookup_bf16_mirror(0, dk, n, out, mirror, device_keys=True)
        else:
            s.lookup_bf16_mirror(0, keys, n, out, mirror)
        assert s.stats().misses > 0
        assert np.array_equal(out.cpu().numpy(), table.lookup(keys))  # fp32 contract unchanged (sync insert forced)
        assert torch.equal(mirror, out.to(torch.bfloat16))
        a = torch.empty((batch, 1), device="cuda")
        b = torch.empty((batch, 1), device="cuda")
        mlp.forward(out, batch, a)
        mlp.forward_bf16(mirror, batch, b)
        torch.cuda.synchronize()
        assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------------
# TF32 mode: against a PURE fp32 model (no operand rounding in the reference).  The reference's dense model is fp32
# (01_model_training.ipynb cells 7,11); TF32 keeps 10 mantissa bits per operand and accumulates in fp32.
# Tolerance: 1e-3 of the output's scale (max |y| of the fp32 model), stated by the verdict of round 1.
# ------------------------------------------------------------------------------------------------
def fp32_model(torch, x, weights, biases, relu):
    h = x.double()
    for l, w in enumerate(weights):
        h = h @ torch.from_numpy(w).cuda().double().t()
        if biases[l] is not None:
            h = h + torch.from_numpy(biases[l]).cuda().double()
        if relu[l]:
            h = torch.relu(h)
    return h


@pytest.mark.parametrize("batch,dims,relu", [
    (128, [32, 128], [0]),                       # one tile, one k-block of 32 fp32
    (128, [256, 128], [0]),                      # ring wrap-around (8 k-blocks, 4 stages)
    (1000, [48, 256, 128, 1], [0, 0, 0]),        # the sample's dense model, ragged batch
    (4096, [3328, 1024, 512, 256, 1], [1, 1, 1, 0]),  # Criteo-shape head
    (333, [136, 200, 72], [1, 0]),               # N and K tails
])
def test_tf32_mlp_is_within_1e3_of_a_pure_fp32_model(cuda_device, batch, dims, relu):
    torch = _torch()
    rng = np.random.default_rng(7 * batch + len(dims))
    weights = [(rng.standard_normal((dims[l + 1], dims[l])) / np.sqrt(dims[l])).astype(np.float32) for l in range(len(dims) - 1)]
    biases = [rng.standard_normal(dims[l + 1]).astype(np.float32) * 0.1 if l % 2 == 0 else None for l in range(len(dims) - 1)]
    mlp = hb.DenseMlp(0, weights, biases, relu, precision="tf32")
    x = torch.from_numpy(rng.standard_normal((batch, dims[0])).astype(np.float32)).cuda()
    out = torch.full((batch, dims[-1]), float("nan"), device="cuda")
    mlp.forward(x, batch, out)
    torch.cuda.synchronize()
    want = fp32_model(torch, x, weights, biases, relu)
    assert torch.isfinite(out).all()
    scale = float(want.abs().max())
    err = float((out.double() - want).abs().max())
    assert err <= 1e-3 * scale, (err, scale)
    # and the bf16 head on the same model, for the record: an order of magnitude further away
    mlp16 = hb.DenseMlp(0, weights, biases, relu)
    out16 = torch.empty_like(out)
    mlp16.forward(x, batch, out16)
    torch.cuda.synchronize()
    err16 = float((out16.double() - want).abs().max())
    assert err < err16
    with pytest.raises(Exception, match="TF32"):
        mlp.forward_bf16(x.to(torch.bfloat16), batch, out)
    mlp.close()
    mlp16.close()
