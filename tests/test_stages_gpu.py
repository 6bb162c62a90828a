"""Each CUDA stage kernel against the same-layout CPU restatement (oracle/forward_laplacian.py, fp64).
Tolerances are fp32 round-off relative to the largest entry of the tensor, written per test."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import forward_laplacian as FL
from oracle import philox as PH
from oracle import psiformer_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from psiformer_torch_b200 import _lib
    return _lib


def _dev(t):
    return t.to(torch.float32).cuda().contiguous()


def _rel(got, ref):
    ref = ref.double()
    return ((got.double().cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _payload(B, N, width, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    Cc = FL.n_channels(N)
    return torch.randn(B, N, Cc, width, generator=g, dtype=torch.float64) * scale


@pytest.mark.parametrize("name", ["debug", "he_small", "be", "lih", "n2"])
def test_embed(golden, lib, name):
    from gpu_util import make_engine
    sysm, params, data = golden(name)
    eng = make_engine(sysm, params)
    x = data["x"]
    ref = FL.embed_payload(sysm, O.cast_params(params, torch.float64), x.double())
    got = eng.stage_embed(x, FL.n_channels(sysm.n_elec), sysm.n_embd)
    assert _rel(got, ref) < 2e-6
    got1 = eng.stage_embed(x, 1, sysm.n_embd)
    assert _rel(got1[:, :, 0], ref[:, :, 0]) < 2e-6


@pytest.mark.parametrize("rows_shape,k_in,n_out", [((3, 2, 8), 4, 12), ((5, 4, 14), 256, 768), ((2, 10, 32), 1024, 256),
                                                   ((7, 3, 11), 64, 160), ((1, 1, 5), 6, 3)])
def test_linear(lib, rows_shape, k_in, n_out):
    B, N, Cc = rows_shape
    g = torch.Generator().manual_seed(k_in + n_out)
    P = torch.randn(B, N, Cc, k_in, generator=g, dtype=torch.float64)
    W = torch.randn(n_out, k_in, generator=g, dtype=torch.float64) / k_in ** 0.5
    b = torch.randn(n_out, generator=g, dtype=torch.float64)
    res = torch.randn(B, N, Cc, n_out, generator=g, dtype=torch.float64)
    ref = FL.linear_payload(P, W, b) + res
    Pd, Wd, bd, rd = _dev(P), _dev(W), _dev(b), _dev(res)
    out = torch.empty(B, N, Cc, n_out, dtype=torch.float32, device="cuda")
    lib.check(lib.load().psif_stage_linear(Pd.data_ptr(), Wd.data_ptr(), bd.data_ptr(), rd.data_ptr(), B * N * Cc, Cc,
                                           k_in, n_out, 0, out.data_ptr(), _stream()))
    assert _rel(out, ref) < 5e-6
    # in-place residual (the pipeline accumulates into the residual stream) and the value-path GELU epilogue
    lib.check(lib.load().psif_stage_linear(Pd.data_ptr(), Wd.data_ptr(), bd.data_ptr(), rd.data_ptr(), B * N * Cc, Cc,
                                           k_in, n_out, 0, rd.data_ptr(), _stream()))
    assert _rel(rd, ref) < 5e-6
    v = torch.nn.functional.gelu(P.reshape(-1, k_in) @ W.t() + b, approximate="tanh")
    out1 = torch.empty(B * N * Cc, n_out, dtype=torch.float32, device="cuda")
    lib.check(lib.load().psif_stage_linear(Pd.data_ptr(), Wd.data_ptr(), bd.data_ptr(), None, B * N * Cc, 1, k_in, n_out,
                                           1, out1.data_ptr(), _stream()))
    assert _rel(out1, v) < 5e-6


@pytest.mark.parametrize("B,N,d", [(3, 3, 4), (4, 2, 64), (3, 10, 256), (2, 14, 256), (2, 4, 96), (2, 2, 512)])
def test_layernorm(lib, B, N, d):
    P = _payload(B, N, d, 3 + d, 1.0) + 0.3
    g = torch.Generator().manual_seed(d)
    gamma = 1 + 0.2 * torch.randn(d, generator=g, dtype=torch.float64)
    beta = 0.1 * torch.randn(d, generator=g, dtype=torch.float64)
    ref = FL.layernorm_payload(P, gamma, beta)
    Pd, gd, bd = _dev(P), _dev(gamma), _dev(beta)
    out = torch.empty_like(Pd)
    Cc = P.shape[2]
    lib.check(lib.load().psif_stage_layernorm(Pd.data_ptr(), gd.data_ptr(), bd.data_ptr(), B * N, Cc, d, 0, out.data_ptr(), _stream()))
    assert _rel(out[:, :, :-1], ref[:, :, :-1]) < 5e-6
    assert _rel(out[:, :, -1], ref[:, :, -1]) < 2e-5      # Laplacian row: a sum of 3N fp32 products
    P1 = Pd[:, :, 0].contiguous()
    out1 = torch.empty_like(P1)
    lib.check(lib.load().psif_stage_layernorm(P1.data_ptr(), gd.data_ptr(), bd.data_ptr(), B * N, 1, d, 0, out1.data_ptr(), _stream()))
    assert _rel(out1, ref[:, :, 0]) < 5e-6


@pytest.mark.parametrize("B,N,d,H", [(3, 3, 4, 2), (4, 2, 64, 16), (3, 10, 256, 4), (2, 14, 256, 4), (2, 6, 256, 32),
                                     (37, 3, 32, 8), (70, 2, 64, 16),
                                     # warp-per-unit kernel (5..14 electrons, head_dim 64), odd counts = zero-padded rows
                                     (5, 5, 256, 4), (3, 7, 256, 4), (2, 9, 128, 2), (3, 13, 256, 4), (2, 12, 256, 4),
                                     (9, 8, 64, 1), (2, 6, 128, 2), (3, 11, 256, 4),
                                     (2, 15, 256, 4), (2, 16, 256, 4)])
def test_attention(lib, B, N, d, H):
    QKV = _payload(B, N, 3 * d, 11 + d + N, 0.7)
    ref = FL.attention_payload(QKV, H)
    Qd = _dev(QKV)
    Cc = QKV.shape[2]
    out = torch.empty(B, N, Cc, d, dtype=torch.float32, device="cuda")
    lib.check(lib.load().psif_stage_attention(Qd.data_ptr(), B, N, Cc, d, H, 0, out.data_ptr(), _stream()))
    assert _rel(out[:, :, :-1], ref[:, :, :-1]) < 1e-5
    assert _rel(out[:, :, -1], ref[:, :, -1]) < 5e-5
    Q1 = Qd[:, :, 0].contiguous()
    out1 = torch.empty(B, N, d, dtype=torch.float32, device="cuda")
    lib.check(lib.load().psif_stage_attention(Q1.data_ptr(), B, N, 1, d, H, 0, out1.data_ptr(), _stream()))
    assert _rel(out1, ref[:, :, 0]) < 1e-5


@pytest.mark.parametrize("B,N,d,H", [(5, 4, 256, 4), (3, 2, 128, 2), (4, 5, 256, 4), (3, 7, 64, 1), (2, 10, 256, 4),
                                     (3, 14, 256, 4), (2, 16, 128, 2), (40, 4, 256, 4)])
def test_attention_first_layer(lib, B, N, d, H):
    """The first layer's attention takes the compact payload [token][5][3 d] (value, own tangents, Laplacian: the only
    non-zero rows in front of the first attention) and must give what the dense rule gives on the zero-expanded payload."""
    g = torch.Generator().manual_seed(100 * N + d)
    q5 = torch.randn(B, N, 5, 3 * d, generator=g, dtype=torch.float64) * 0.7
    Cc = 3 * N + 2
    dense = torch.zeros(B, N, Cc, 3 * d, dtype=torch.float64)
    dense[:, :, 0] = q5[:, :, 0]
    dense[:, :, -1] = q5[:, :, 4]
    for i in range(N):
        dense[:, i, 1 + 3 * i:4 + 3 * i] = q5[:, i, 1:4]
    ref = FL.attention_payload(dense, H)
    out = torch.empty(B, N, Cc, d, dtype=torch.float32, device="cuda")
    lib.check(lib.load().psif_stage_attention_first_layer(_dev(q5).data_ptr(), B, N, d, H, 0, out.data_ptr(), _stream()))
    assert _rel(out[:, :, :-1], ref[:, :, :-1]) < 1e-5
    assert _rel(out[:, :, -1], ref[:, :, -1]) < 5e-5
    # packed output: the same values as the fp16 pair
    pk = torch.empty(B, N, Cc, d, dtype=torch.float32, device="cuda")
    lib.check(lib.load().psif_stage_attention_first_layer(_dev(q5).data_ptr(), B, N, d, H, 1, pk.data_ptr(), _stream()))
    ref_pk = torch.empty_like(pk)
    lib.check(lib.load().psif_stage_pack(out.data_ptr(), B * N * Cc, d, ref_pk.data_ptr(), None, _stream()))
    torch.cuda.synchronize()
    assert torch.equal(pk.view(torch.int32), ref_pk.view(torch.int32))


@pytest.mark.parametrize("B,N,w", [(3, 3, 16), (2, 10, 1024), (2, 14, 1024)])
def test_gelu(lib, B, N, w):
    P = _payload(B, N, w, 5 + w + N, 1.5)
    ref = FL.gelu_payload(P)
    Pd = _dev(P)
    out = torch.empty_like(Pd)
    lib.check(lib.load().psif_stage_gelu(Pd.data_ptr(), B * N, P.shape[2], w, out.data_ptr(), _stream()))
    assert _rel(out, ref) < 1e-5
    lib.check(lib.load().psif_stage_gelu(Pd.data_ptr(), B * N, P.shape[2], w, Pd.data_ptr(), _stream()))   # in place
    assert _rel(Pd, ref) < 1e-5


def test_slogdet_known_answers_and_random(lib):
    import os
    from conftest import GOLDEN_DIR
    from psiformer_torch_b200.logdet_matmul import logdet_matmul
    z = np.load(os.path.join(GOLDEN_DIR, "logdet_kat.npz"))
    la, sg = logdet_matmul(torch.from_numpy(z["x1"]).float().cuda(), torch.from_numpy(z["x2"]).float().cuda(),
                           torch.from_numpy(z["w"]).float().cuda())
    # reference fp64 on the fp32-rounded inputs
    rl, rs = O.logdet_matmul_value(torch.from_numpy(z["x1"]).float().double(), torch.from_numpy(z["x2"]).float().double(),
                                   torch.from_numpy(z["w"]).float().double())
    assert torch.allclose(la.double().cpu(), rl, rtol=1e-5, atol=1e-5)
    assert torch.equal(sg.double().cpu(), rs)
    la2, sg2 = logdet_matmul(torch.from_numpy(z["near"]).float().cuda(), torch.ones(1, 1, 1, 1, device="cuda"),
                             torch.ones(1, 1, device="cuda"))
    r2, s2 = O.logdet_matmul_value(torch.from_numpy(z["near"]).float().double(), torch.ones(1, 1, 1, 1, dtype=torch.float64),
                                   torch.ones(1, 1, dtype=torch.float64))
    assert torch.allclose(la2.double().cpu(), r2, rtol=1e-5, atol=1e-5) and torch.equal(sg2.double().cpu(), s2)
    g = torch.Generator().manual_seed(9)
    for K, nu, nd in [(1, 1, 1), (4, 4, 2), (16, 5, 5), (32, 7, 7), (3, 8, 6), (64, 2, 2)]:
        x1 = torch.randn(50, K, nu, nu, generator=g)
        x2 = torch.randn(50, K, nd, nd, generator=g)
        x1[0, 0] = 0.0                       # singular block: exercises the exact 1e-6 clamp path
        if nu > 1:
            x1[1, :, 1] = x1[1, :, 0]        # rank deficient in every determinant
        w = torch.softmax(torch.randn(K, 2, generator=g), 0)
        la, sg = logdet_matmul(x1.cuda(), x2.cuda(), w.cuda())
        rl, rs = O.logdet_matmul_value(x1.double(), x2.double(), w.double())
        assert torch.allclose(la.double().cpu(), rl, rtol=2e-5, atol=2e-5), (K, nu, nd, (la.double().cpu() - rl).abs().max())
        assert torch.equal(sg.double().cpu(), rs), (K, nu, nd)
    with pytest.raises(ValueError):
        logdet_matmul(torch.zeros(2, 3, 2, 2, device="cuda"), torch.zeros(3, 3, 2, 2, device="cuda"), torch.ones(3, 1, device="cuda"))
    with pytest.raises(ValueError):
        logdet_matmul(torch.zeros(2, 3, 2, 2, device="cuda"), torch.zeros(2, 3, 2, 2, device="cuda"), torch.ones(4, 1, device="cuda"))


@pytest.mark.parametrize("name", ["he_small", "large", "ne", "n2"])
def test_jastrow_and_potential(golden, name):
    from psiformer_torch_b200.hamiltonian import Potential
    from psiformer_torch_b200.jastrow import Jastrow
    sysm, params, data = golden(name)
    x = data["x"]
    j = Jastrow(sysm.n_up, sysm.n_dn).cuda()
    with torch.no_grad():
        j.alpha_anti.copy_(params["jastrow.alpha_anti"])
        j.alpha_par.copy_(params["jastrow.alpha_par"])
    ref = O.jastrow(sysm, O.cast_params(params, torch.float64), x.double())
    assert torch.allclose(j(x.cuda()).double().cpu(), ref, rtol=1e-6, atol=1e-6)
    v = Potential(x.cuda(), nuclei=sysm.nuclei).potential()
    assert torch.allclose(v.double().cpu(), O.potential(sysm, x.double()), rtol=1e-6, atol=1e-5)
    with pytest.raises(ValueError):
        j(torch.zeros(3, 1, 3, device="cuda"))


def test_philox_stream_matches_numpy_oracle(lib):
    assert PH.known_answer()
    for seed, w0, step, B, N in [(7, 0, 0, 1000, 4), (2**40 + 3, 123456789012, 2**33 + 5, 257, 10)]:
        nrm = torch.empty(B, N, 3, dtype=torch.float32, device="cuda")
        uni = torch.empty(B, dtype=torch.float32, device="cuda")
        lib.check(lib.load().psif_philox_normal(seed, w0, step, B, N, nrm.data_ptr(), uni.data_ptr(), _stream()))
        assert np.array_equal(uni.cpu().numpy(), PH.mh_uniforms(seed, w0, step, B))      # integer path: bit-exact
        assert np.allclose(nrm.cpu().numpy(), PH.mh_normals(seed, w0, step, B, N), rtol=0, atol=2e-6)


@pytest.mark.parametrize("rows,k_in,n_out,Cc", [(640, 256, 256, 1), (4096 + 77, 256, 768, 14), (2048, 1024, 256, 32),
                                                  (20000, 256, 1024, 14), (4096, 256, 64, 14), (3000, 256, 160, 32)])
@pytest.mark.parametrize("variant", [0, 1])
def test_linear_tcgen05_3xtf32(lib, rows, k_in, n_out, Cc, variant):
    """The tensor-core Linear (tcgen05.mma, three-pass split: gemm mode 0 = fp16 halves with the low half scaled by
    2^11, mode 1 = tf32 halves) against fp64, and against the exact-fp32 FFMA kernel: the split GEMM must stay within
    a small multiple of plain fp32 round-off."""
    _linear_tc_case(lib, rows, k_in, n_out, Cc, variant)


def _linear_tc_case(lib, rows, k_in, n_out, Cc, variant):
    MODE = variant
    g = torch.Generator().manual_seed(rows + n_out)
    X = torch.randn(rows, k_in, generator=g, dtype=torch.float64)
    W = torch.randn(n_out, k_in, generator=g, dtype=torch.float64) / k_in ** 0.5
    b = torch.randn(n_out, generator=g, dtype=torch.float64)
    res = torch.randn(rows, n_out, generator=g, dtype=torch.float64)
    ref = X.float().double() @ W.float().double().t()
    ref[torch.arange(rows) % Cc == 0] += b.float().double()
    ref = ref + res.float().double()
    Xd, Wd, bd, rd = _dev(X), _dev(W), _dev(b), _dev(res)
    scratch = torch.empty(3 * n_out * k_in + 4, dtype=torch.float32, device="cuda")
    out = torch.empty(rows, n_out, dtype=torch.float32, device="cuda")
    L = lib.load()
    lib.check(L.psif_stage_linear_tc(Xd.data_ptr(), Wd.data_ptr(), bd.data_ptr(), rd.data_ptr(), rows, Cc, k_in, n_out, 0, MODE, 0,
                                     out.data_ptr(), scratch.data_ptr(), None, _stream()))
    out_f = torch.empty_like(out)
    lib.check(L.psif_stage_linear(Xd.data_ptr(), Wd.data_ptr(), bd.data_ptr(), rd.data_ptr(), rows, Cc, k_in, n_out, 0,
                                  out_f.data_ptr(), _stream()))
    torch.cuda.synchronize()
    e_tc, e_ff = _rel(out, ref), _rel(out_f, ref)
    print(f"\n[tcgen05 variant {variant} {rows}x{k_in}x{n_out}] rel err split GEMM {e_tc:.2e}  FFMA {e_ff:.2e}")
    assert e_tc < 2e-6 and e_tc < 8 * e_ff + 2e-7
    # in-place residual + GELU epilogue (value path)
    lib.check(L.psif_stage_linear_tc(Xd.data_ptr(), Wd.data_ptr(), bd.data_ptr(), rd.data_ptr(), rows, Cc, k_in, n_out, 0, MODE, 0,
                                     rd.data_ptr(), scratch.data_ptr(), None, _stream()))
    assert torch.equal(rd, out)
    out_g = torch.empty_like(out)
    lib.check(L.psif_stage_linear_tc(Xd.data_ptr(), Wd.data_ptr(), bd.data_ptr(), None, rows, 1, k_in, n_out, 1, MODE, 0,
                                     out_g.data_ptr(), scratch.data_ptr(), None, _stream()))
    refg = torch.nn.functional.gelu(X.float().double() @ W.float().double().t() + b.float().double(), approximate="tanh")
    assert _rel(out_g, refg) < 2e-6


@pytest.mark.parametrize("tokens,Cc,k_in,n_out", [(300, 14, 256, 1024), (1171, 14, 256, 512), (256, 8, 128, 512),
                                                  (100, 32, 256, 1024), (64, 44, 256, 1024), (147, 11, 64, 256),
                                                  (700, 1, 256, 1024)])
def test_linear_tcgen05_fused_payload_gelu(lib, tokens, Cc, k_in, n_out):
    """MLP up-projection with the payload GELU applied by the GEMM epilogue (token-aligned row tiles) against the
    un-fused Linear followed by the GELU payload kernel, which the other stage tests pin to the oracle.  Value and
    tangent rows are bit-identical; the Laplacian row sums the squared tangents in a different order."""
    rows = tokens * Cc
    MODE = 0
    g = torch.Generator().manual_seed(tokens + Cc)
    Xd = torch.randn(rows, k_in, generator=g).cuda()
    Wd = (torch.randn(n_out, k_in, generator=g) / k_in ** 0.5).cuda()
    bd = torch.randn(n_out, generator=g).cuda()
    scratch = torch.empty(3 * n_out * k_in + 4, dtype=torch.float32, device="cuda")
    L = lib.load()
    plain = torch.empty(rows, n_out, dtype=torch.float32, device="cuda")
    lib.check(L.psif_stage_linear_tc(Xd.data_ptr(), Wd.data_ptr(), bd.data_ptr(), None, rows, Cc, k_in, n_out, 0, MODE, 0,
                                     plain.data_ptr(), scratch.data_ptr(), None, _stream()))
    want = torch.empty_like(plain)
    lib.check(L.psif_stage_gelu(plain.data_ptr(), tokens, Cc, n_out, want.data_ptr(), _stream()))
    got = torch.full_like(plain, float("nan"))
    lib.check(L.psif_stage_linear_tc(Xd.data_ptr(), Wd.data_ptr(), bd.data_ptr(), None, rows, Cc, k_in, n_out, 2, MODE, 0,
                                     got.data_ptr(), scratch.data_ptr(), None, _stream()))
    torch.cuda.synchronize()
    g3, w3 = got.view(tokens, Cc, n_out), want.view(tokens, Cc, n_out)
    assert torch.equal(g3[:, :max(Cc - 1, 1)], w3[:, :max(Cc - 1, 1)])
    err = (g3[:, -1] - w3[:, -1]).abs().max().item()
    assert err <= 4e-6 * w3[:, -1].abs().max().item(), err


def _unpack(t, width):
    """packed fp16 pair rows (psif_stage_pack layout) -> (h0, h1) as float64 and their recombination."""
    h = t.view(torch.float16).view(t.shape[0], 2 * width)
    h0, h1 = h[:, :width].double(), h[:, width:].double()
    return h0, h1, h0 + h1 / 2048.0


def test_packed_pair_format_and_range_flag(lib):
    L = lib.load()
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(300, 256, generator=g) * torch.logspace(-6, 4, 300)[:, None]).cuda()
    out = torch.empty_like(x)
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    lib.check(L.psif_stage_pack(x.data_ptr(), 300, 256, out.data_ptr(), flag.data_ptr(), _stream()))
    h0, h1, rec = _unpack(out, 256)
    assert torch.equal(h0.float(), x.half().float())                       # h0 = fp16(x)
    err = (rec - x.double()).abs()
    assert (err <= 2.0 ** -22 * x.double().abs() + 2.0 ** -36).all()       # oracle/fp16_split.py bound
    assert int(flag.item()) == 0
    x[5, 7] = 7e4
    lib.check(L.psif_stage_pack(x.data_ptr(), 300, 256, out.data_ptr(), flag.data_ptr(), _stream()))
    assert int(flag.item()) == 1


@pytest.mark.parametrize("Cc,N,H,d", [(14, 4, 4, 256), (1, 4, 4, 256), (32, 10, 4, 256), (44, 14, 4, 256), (23, 7, 2, 128)])
def test_producers_write_the_same_pair_as_the_pack_pass(lib, Cc, N, H, d):
    """LayerNorm / attention with packed output == pack(fp32 output), bit for bit."""
    L = lib.load()
    g = torch.Generator().manual_seed(Cc)
    B = 37
    P = torch.randn(B * N * Cc, d, generator=g).cuda()
    gam, bet = torch.randn(d, generator=g).cuda(), torch.randn(d, generator=g).cuda()
    f32, pk, ref = torch.empty_like(P), torch.empty_like(P), torch.empty_like(P)
    lib.check(L.psif_stage_layernorm(P.data_ptr(), gam.data_ptr(), bet.data_ptr(), B * N, Cc, d, 0, f32.data_ptr(), _stream()))
    lib.check(L.psif_stage_layernorm(P.data_ptr(), gam.data_ptr(), bet.data_ptr(), B * N, Cc, d, 1, pk.data_ptr(), _stream()))
    lib.check(L.psif_stage_pack(f32.data_ptr(), P.shape[0], d, ref.data_ptr(), None, _stream()))
    assert torch.equal(pk.view(torch.int32), ref.view(torch.int32))
    Q = torch.randn(B * N * Cc, 3 * d, generator=g).cuda()
    lib.check(L.psif_stage_attention(Q.data_ptr(), B, N, Cc, d, H, 0, f32.data_ptr(), _stream()))
    lib.check(L.psif_stage_attention(Q.data_ptr(), B, N, Cc, d, H, 1, pk.data_ptr(), _stream()))
    lib.check(L.psif_stage_pack(f32.data_ptr(), P.shape[0], d, ref.data_ptr(), None, _stream()))
    assert torch.equal(pk.view(torch.int32), ref.view(torch.int32))


@pytest.mark.parametrize("tokens,Cc,k_in,n_out", [(300, 14, 256, 1024), (1171, 14, 256, 768), (100, 32, 1024, 256), (64, 44, 256, 1024),
                                                  (2000, 14, 256, 64)])
def test_linear_tcgen05_packed_operand(lib, tokens, Cc, k_in, n_out):
    """The Linear on an A operand that arrives as the packed fp16 pair (no splitter) == the Linear that splits fp32 rows
    itself, bit for bit, incl. K passes, the ragged orbital-head tile, and the fused payload GELU with packed output."""
    L = lib.load()
    rows = tokens * Cc
    g = torch.Generator().manual_seed(tokens)
    X = torch.randn(rows, k_in, generator=g).cuda()
    W = (torch.randn(n_out, k_in, generator=g) / k_in ** 0.5).cuda()
    b = torch.randn(n_out, generator=g).cuda()
    res = torch.randn(rows, n_out, generator=g).cuda()
    scratch = torch.empty(3 * n_out * k_in + 4, dtype=torch.float32, device="cuda")
    Xp = torch.empty_like(X)
    lib.check(L.psif_stage_pack(X.data_ptr(), rows, k_in, Xp.data_ptr(), None, _stream()))
    a, c = res.clone(), res.clone()
    lib.check(L.psif_stage_linear_tc(X.data_ptr(), W.data_ptr(), b.data_ptr(), a.data_ptr(), rows, Cc, k_in, n_out, 0, 0, 0,
                                     a.data_ptr(), scratch.data_ptr(), None, _stream()))
    lib.check(L.psif_stage_linear_tc(Xp.data_ptr(), W.data_ptr(), b.data_ptr(), c.data_ptr(), rows, Cc, k_in, n_out, 0, 0, 1,
                                     c.data_ptr(), scratch.data_ptr(), None, _stream()))
    if k_in <= 512:
        assert torch.equal(a, c)
    else:      # two K passes: the packed-operand kernel sums both halves in its epilogue, res + (p0 + p1) instead of (res + p0) + p1
        assert (a - c).abs().max().item() <= 1e-6 * a.abs().max().item()
    if n_out % 128 == 0 and k_in <= 512:
        f32 = torch.empty(rows, n_out, device="cuda")
        pk, ref = torch.empty_like(f32), torch.empty_like(f32)
        lib.check(L.psif_stage_linear_tc(X.data_ptr(), W.data_ptr(), b.data_ptr(), None, rows, Cc, k_in, n_out, 2, 0, 0,
                                         f32.data_ptr(), scratch.data_ptr(), None, _stream()))
        lib.check(L.psif_stage_linear_tc(Xp.data_ptr(), W.data_ptr(), b.data_ptr(), None, rows, Cc, k_in, n_out, 2, 0, 1,
                                         pk.data_ptr(), scratch.data_ptr(), None, _stream()))
        lib.check(L.psif_stage_pack(f32.data_ptr(), rows, n_out, ref.data_ptr(), None, _stream()))
        # value and tangent rows bit for bit; the Laplacian row sums the squared tangents in a different order in the
        # packed-operand kernel (eight row classes per warp instead of four)
        p3, r3 = pk.view(torch.int32).view(tokens, Cc, n_out), ref.view(torch.int32).view(tokens, Cc, n_out)
        assert torch.equal(p3[:, :Cc - 1], r3[:, :Cc - 1])
        lap_pk = _unpack(pk.view(tokens, Cc, n_out)[:, -1].contiguous(), n_out)[2]
        lap_f32 = f32.view(tokens, Cc, n_out)[:, -1].double()
        assert (lap_pk - lap_f32).abs().max().item() <= 4e-6 * lap_f32.abs().max().item()


@pytest.mark.parametrize("rows,k_in,n_out", [(16384, 256, 1024), (1000, 128, 512)])
def test_linear_tcgen05_packed_plain_gelu_writes_the_pair(lib, rows, k_in, n_out):
    """Value path (plain rows): MLP up-projection with a packed A operand and the GELU in the epilogue writes its output
    as the packed pair, bit for bit pack(fp32 output of the splitting kernel)."""
    L = lib.load()
    g = torch.Generator().manual_seed(rows)
    X = torch.randn(rows, k_in, generator=g).cuda()
    W = (torch.randn(n_out, k_in, generator=g) / k_in ** 0.5).cuda()
    b = torch.randn(n_out, generator=g).cuda()
    scratch = torch.empty(3 * n_out * k_in + 4, dtype=torch.float32, device="cuda")
    Xp = torch.empty_like(X)
    lib.check(L.psif_stage_pack(X.data_ptr(), rows, k_in, Xp.data_ptr(), None, _stream()))
    f32 = torch.empty(rows, n_out, device="cuda")
    pk, ref = torch.full_like(f32, float("nan")), torch.empty_like(f32)
    lib.check(L.psif_stage_linear_tc(X.data_ptr(), W.data_ptr(), b.data_ptr(), None, rows, 1, k_in, n_out, 1, 0, 0,
                                     f32.data_ptr(), scratch.data_ptr(), None, _stream()))
    lib.check(L.psif_stage_linear_tc(Xp.data_ptr(), W.data_ptr(), b.data_ptr(), None, rows, 1, k_in, n_out, 1, 0, 1,
                                     pk.data_ptr(), scratch.data_ptr(), None, _stream()))
    lib.check(L.psif_stage_pack(f32.data_ptr(), rows, n_out, ref.data_ptr(), None, _stream()))
    torch.cuda.synchronize()
    assert torch.equal(pk.view(torch.int32), ref.view(torch.int32))


def test_packed_pipeline_matches_the_splitting_one(golden, monkeypatch):
    """End to end: producers writing the pair (default) vs GEMMs splitting fp32 activations (PSIF_PACK_PRODUCERS=0)."""
    from gpu_util import make_engine
    sysm, params, data = golden("be")
    x = data["x"].cuda()
    a = make_engine(sysm, params).local_energy(x, want_grad=True)
    monkeypatch.setenv("PSIF_PACK_PRODUCERS", "0")
    b = make_engine(sysm, params).local_energy(x, want_grad=True)
    # same operands everywhere; the packed-operand kernels differ in summation ORDER only (the GELU epilogue's sum of squared
    # tangents; the down-projection adds its two K halves before the residual instead of after): last-bit differences
    assert ((a["logabs"] - b["logabs"]).abs() <= 2e-6 * a["logabs"].abs().clamp_min(1.0)).all()
    assert (a["grad"] - b["grad"]).abs().max().item() <= 2e-5 * a["grad"].abs().max().item()
    assert (a["e_loc"] - b["e_loc"]).abs().max().item() < 5e-5


@pytest.mark.parametrize("name", ["be", "lih", "ne", "n2"])
def test_first_layer_compact_payload_matches_the_dense_one(golden, monkeypatch, name):
    """In front of the first attention a token's payload has five non-zero rows (value, own tangents, Laplacian); the
    pipeline keeps only those through the first LayerNorm / QKV GEMM and the first attention works from them.  The rows it
    drops are exact zeros in the dense pipeline (PSIF_L0_SPARSE=0)."""
    from gpu_util import make_engine
    from oracle import psiformer_oracle as O
    sysm, params, data = golden(name)
    x = torch.cat([data["x"], O.synthetic_walkers(sysm, 333, 5)]).cuda()        # enough rows for the tensor-core path
    a = make_engine(sysm, params).local_energy(x, want_grad=True)
    monkeypatch.setenv("PSIF_L0_SPARSE", "0")
    b = make_engine(sysm, params).local_energy(x, want_grad=True)
    # LayerNorm and the QKV GEMM give the same bits for the rows they keep; the first attention (attention_first_layer.cuh)
    # sums in another order.  Both pipelines are rounding-level approximations of the fp64 oracle, and ill-conditioned
    # synthetic walkers amplify the rounding: the compact pipeline must be as close to the oracle as the dense one
    ref = O.log_psi(sysm, O.cast_params(params, torch.float64), x.cpu().double())
    ea = (a["logabs"].cpu().double() - ref).abs()
    eb = (b["logabs"].cpu().double() - ref).abs()
    for qt, fac in ((0.5, 1.5), (0.9, 1.5), (0.99, 3.0)):          # the 99th percentile is the 4th worst walker: noisy
        assert ea.quantile(qt).item() <= fac * eb.quantile(qt).item() + 2e-6, (qt, ea.quantile(qt).item(), eb.quantile(qt).item())
    rel = ((a["e_loc"] - b["e_loc"]).abs() / b["e_loc"].abs().clamp_min(1.0)).double()
    assert rel.median().item() < 5e-6 and rel.quantile(0.9).item() < 2e-4, (rel.median().item(), rel.quantile(0.9).item())
    gerr = (a["grad"] - b["grad"]).flatten(1).norm(dim=1) / b["grad"].flatten(1).norm(dim=1).clamp_min(1e-3)
    assert gerr.median().item() < 1e-5 and gerr.quantile(0.9).item() < 1e-3, (gerr.median().item(), gerr.quantile(0.9).item())
    if sysm.n_up + sysm.n_dn == 4:
        monkeypatch.setenv("PSIF_L0_SPARSE", "1")
        monkeypatch.setenv("PSIF_L0_N4", "1")          # the dense 4-electron kernel on zero-filled staging: bit-identical
        c = make_engine(sysm, params).local_energy(x, want_grad=True)
        for k in ("logabs", "e_loc", "grad"):
            assert torch.equal(c[k], b[k]), k


def test_clamp_active_determinant_derivatives_follow_the_reference(lib):
    """Walkers on which the reference's 1e-6 singular-value clamp (logdet_matmul.py:50-51) is ACTIVE: value, gradient and
    Laplacian of log|sum_k w_k det det| must be those of the clamped function, i.e. what torch autograd gives through the
    oracle's svd -> clamp -> log (the clamped direction carries no gradient, the remaining singular values their
    second-order perturbation).  Orbital blocks A_k(t) = A_k + sum_c t_c E_kc + 1/2 sum_c t_c^2 F_kc with one block per
    walker built to have a singular value of 1e-8 .. 3e-7; the payload's tangent channels are E, its Laplacian channel
    sum_c F.  Without the fix-up kernel the same call returns the unclamped derivatives (checked to differ)."""
    L = lib.load()
    torch.manual_seed(11)
    B, K, n_up, n_dn, T = 6, 3, 3, 2, 5
    N, C, Korb = n_up + n_dn, T + 2, K * (n_up + n_dn)
    dd = torch.float64

    def blocks(n):
        a = torch.randn(B, K, n, n, dtype=dd) * 0.7
        return a

    A_up, A_dn = blocks(n_up), blocks(n_dn)
    # one clamp-active block per walker (alternating spin): A + 1e-4 I = U diag(s) V^T with s_min tiny
    for b in range(B):
        n = n_up if b % 2 == 0 else n_dn
        tgt = A_up if b % 2 == 0 else A_dn
        u, _ = torch.linalg.qr(torch.randn(n, n, dtype=dd))
        v, _ = torch.linalg.qr(torch.randn(n, n, dtype=dd))
        s = torch.linspace(1.5, 0.5, n, dtype=dd)
        s[-1] = [1e-8, 3e-7, 5e-8, 1e-7, 2e-8, 2e-7][b]
        tgt[b, b % K] = (u * s) @ v.T - 1e-4 * torch.eye(n, dtype=dd)
    E_up, E_dn = torch.randn(T, B, K, n_up, n_up, dtype=dd) * 0.3, torch.randn(T, B, K, n_dn, n_dn, dtype=dd) * 0.3
    F_up, F_dn = torch.randn(T, B, K, n_up, n_up, dtype=dd) * 0.2, torch.randn(T, B, K, n_dn, n_dn, dtype=dd) * 0.2
    w = torch.tensor([0.9, -0.4, 0.6], dtype=dd)
    # everything the kernel sees is fp32: round the inputs first so that oracle and kernel differentiate the same function
    A_up, A_dn, E_up, E_dn, F_up, F_dn, w = (t.float().double() for t in (A_up, A_dn, E_up, E_dn, F_up, F_dn, w))
    lapA_up, lapA_dn = F_up.sum(0).float().double(), F_dn.sum(0).float().double()

    def f(t):       # (B, T) -> (B,)
        xu = A_up + torch.einsum("bc,cbkij->bkij", t, E_up) + 0.5 * torch.einsum("bc,cbkij->bkij", t * t, F_up)
        xd = A_dn + torch.einsum("bc,cbkij->bkij", t, E_dn) + 0.5 * torch.einsum("bc,cbkij->bkij", t * t, F_dn)
        return O.logdet_matmul_value(xu, xd, w[:, None])[0][:, 0]

    # the Laplacian channel of the payload is sum_c F_c rounded to fp32 as a whole; give the oracle the same second-order term
    t0 = torch.zeros(B, T, dtype=dd, requires_grad=True)
    val = f(t0)
    (g,) = torch.autograd.grad(val.sum(), t0, create_graph=True)
    lap = torch.zeros(B, dtype=dd)
    for c in range(T):
        (h,) = torch.autograd.grad(g[:, c].sum(), t0, retain_graph=True)
        lap = lap + h[:, c]
    # replace the oracle's own sum_c <G, F_c> (from the t^2 term) by <G, fp32(sum_c F_c)>: identical up to the rounding of
    # the channel, which the kernel cannot see through
    smin = O.min_singular_values(A_up, A_dn)
    assert (smin < 1e-6).all()

    phi = torch.zeros(B, N, C, Korb, dtype=torch.float32)
    for k in range(K):
        phi[:, :n_up, 0, k * n_up:(k + 1) * n_up] = A_up[:, k].float()
        phi[:, n_up:, 0, K * n_up + k * n_dn:K * n_up + (k + 1) * n_dn] = A_dn[:, k].float()
        for c in range(T):
            phi[:, :n_up, 1 + c, k * n_up:(k + 1) * n_up] = E_up[c, :, k].float()
            phi[:, n_up:, 1 + c, K * n_up + k * n_dn:K * n_up + (k + 1) * n_dn] = E_dn[c, :, k].float()
        phi[:, :n_up, C - 1, k * n_up:(k + 1) * n_up] = lapA_up[:, k].float()
        phi[:, n_up:, C - 1, K * n_up + k * n_dn:K * n_up + (k + 1) * n_dn] = lapA_dn[:, k].float()
    phi_d, w_d = phi.cuda().contiguous(), w.float().cuda()

    def run(fixup):
        e = torch.empty(B, device="cuda"); la = torch.empty(B, device="cuda"); sg = torch.empty(B, device="cuda")
        gr = torch.empty(B, T, device="cuda"); lp = torch.empty(B, device="cuda")
        st = torch.zeros(B, dtype=torch.int32, device="cuda")
        lib.check(L.psif_stage_det_energy(phi_d.data_ptr(), w_d.data_ptr(), B, N, n_up, C, K, fixup, e.data_ptr(), la.data_ptr(),
                                          sg.data_ptr(), gr.data_ptr(), lp.data_ptr(), st.data_ptr(), _stream()))
        torch.cuda.synchronize()
        return e.double().cpu(), la.double().cpu(), gr.double().cpu(), lp.double().cpu(), st.cpu()

    e1, la1, g1, lp1, st1 = run(1)
    e0, la0, g0, lp0, st0 = run(0)
    assert ((st1 & 2) != 0).all() and ((st0 & 2) != 0).all()          # PSIF_ST_CLAMP_SUSPECT on every walker
    gref, lref = g.detach(), lap.detach()
    assert torch.allclose(la1, val.detach(), rtol=1e-6, atol=1e-6)
    gscale = gref.abs().max().item()
    assert (g1 - gref).abs().max().item() < 2e-5 * max(1.0, gscale), (g1 - gref).abs().max().item()
    lscale = lref.abs().max().item()
    assert (lp1 - lref).abs().max().item() < 1e-4 * max(1.0, lscale), ((lp1 - lref).abs().max().item(), lscale)
    eref = -0.5 * (lref + (gref * gref).sum(1))
    assert (e1 - eref).abs().max().item() < 1e-4 * max(1.0, eref.abs().max().item())
    # and the smooth (unclamped) formulas are far off on these walkers: 1 / s_min sized terms
    assert (g0 - gref).abs().max().item() > 1e3 * (g1 - gref).abs().max().item()
