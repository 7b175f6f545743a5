"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/exvae_b200.h declares, rejects bad arguments without touching a GPU, and the host-side
mirror keeps the reference's state_dict keys."""
import ctypes
import os

import numpy as np

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from build_lib import load_build_module
    B = load_build_module()
    if not os.path.exists(B.LIB):
        B.build()
    from exemplar_vae_b200._lib import lib
    return lib()


def test_library_exports_every_declared_symbol(L):
    from exemplar_vae_b200._lib import LIB_PATH, parse_header
    protos = parse_header()
    assert len(protos) >= 40
    dll = ctypes.CDLL(LIB_PATH)
    for name in protos:
        assert hasattr(dll, name), name
    assert L.exvae_abi_version() == 1
    assert L.exvae_error_string(0) == b"ok"
    assert b"workspace" in L.exvae_error_string(-3)


def test_argument_errors_without_gpu(L):
    # null pointers / non-positive sizes are rejected before any CUDA call
    assert L.exvae_pairwise_distance(None, None, 4, 4, 4, None, None) == -1
    assert L.exvae_gated_dense_fwd(None, None, None, None, None, 1, 1, 1, None, None, None, 0, None) == -1
    assert L.exvae_prior_lse_workspace_bytes(0, 10, 40) == 0
    n = L.exvae_prior_lse_workspace_bytes(512, 25000, 40)
    assert 4e6 < n < 2e8
    # fused K2: per-split candidate lists only, the [B,N] distance matrix never reaches HBM
    assert 0 < L.exvae_knn_workspace_bytes(100, 25000, 40, 10) < 100 * 25000 * 4
    assert L.exvae_gated_dense_bwd_workspace_bytes(25000, 784, 300) > 25000 * 600 * 4
    # the deferred-finish switch is plain host state: returns the previous setting; a flush with nothing queued is a no-op
    assert L.exvae_dense_bwd_defer_finish(1) == 0
    assert L.exvae_dense_bwd_defer_finish(0) == 1
    assert L.exvae_dense_bwd_flush(None) == 0


def test_header_cites_reference_for_each_group():
    src = open(os.path.join(ROOT, "include", "exvae_b200.h")).read()
    for cite in ("utils/distributions.py:12-18", "models/BaseModel.py:98-109", "models/BaseModel.py:263-264",
                 "utils/nn.py:44-69", "utils/optimizer.py:32-80", "utils/knn_on_latent.py:4-9"):
        assert cite in src


def test_state_dict_keys_match_reference_checkpoint_layout():
    import exemplar_vae_b200 as E
    from oracle import exvae_oracle as O
    for name, n_tensors, n_params in (("vae", 23, 1116865), ("hvae_2level", 55, 2431025)):
        args = O.make_args(model_name=name)
        m = E.importing_model(args)(args)
        sd = m.state_dict()
        assert len(sd) == n_tensors and sum(v.numel() for v in sd.values()) == n_params
        ref = O.init_params(args)
        assert set(sd.keys()) == set(ref.keys())
        for k in sd:
            assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    ck = "/root/reference/pretrained_model/exemplar_prior_on_dynamic_mnist_model_name=vae/1/checkpoint_best.pth"
    if os.path.exists(ck):   # build container only
        args = O.make_args(model_name="vae")
        m = E.importing_model(args)(args)
        m.load_state_dict(torch.load(ck, map_location="cpu")["state_dict"])
        assert abs(float(m.prior_log_variance.detach()) + 2.4189) < 1e-3


def test_product_has_no_cpu_path():
    import exemplar_vae_b200 as E
    from oracle import exvae_oracle as O
    args = O.make_args(model_name="vae", hidden_size=16)
    m = E.importing_model(args)(args)
    x = torch.rand(4, 784)
    with pytest.raises(E.ExvaeError):
        m.q_z(x)


def test_set_beta_schedule():
    from types import SimpleNamespace
    import exemplar_vae_b200 as E
    a = SimpleNamespace(warmup=100)
    assert E.set_beta(a, 0) == 0.0 and E.set_beta(a, 50) == 0.5 and E.set_beta(a, 500) == 1.0
    assert E.set_beta(SimpleNamespace(warmup=0), 3) == 1.0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "exemplar_vae_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("against the oracle", ""), fn


def test_row_block_views_autograd_cpu():
    """ops.split_rows / ops.shared_rows are pure autograd plumbing (no kernel): same values and gradients as slicing."""
    import torch
    from exemplar_vae_b200 import ops
    g = torch.Generator().manual_seed(0)
    t0 = torch.randn(50, 12, generator=g)
    w1, w2, wf = torch.randn(20, 12, generator=g), torch.randn(30, 12, generator=g), torch.randn(50, 12, generator=g)
    a = t0.clone().requires_grad_(True)
    x, y = ops.split_rows(a * 1.0, 20)
    assert torch.equal(x, t0[:20]) and torch.equal(y, t0[20:])
    ((x * w1).sum() + (y * w2).sum()).backward()
    b = t0.clone().requires_grad_(True)
    ((b[:20] * w1).sum() + (b[20:] * w2).sum()).backward()
    assert torch.equal(a.grad, b.grad)
    a = t0.clone().requires_grad_(True)
    full, head = ops.shared_rows(a * 1.0, 20)
    ((full * wf).sum() + (head * w1).sum()).backward()
    b = t0.clone().requires_grad_(True)
    ((b * wf).sum() + (b[:20] * w1).sum()).backward()
    assert torch.allclose(a.grad, b.grad, rtol=1e-6, atol=0)
    a = t0.clone().requires_grad_(True)
    x, y = ops.split_rows(a * 1.0, 20)
    (y * w2).sum().backward()                      # only one consumer: the other block's gradient is zero
    assert torch.equal(a.grad[:20], torch.zeros(20, 12)) and torch.equal(a.grad[20:], w2)


def test_workspace_sizes_cover_the_tensor_core_plans(L):
    """Host-side planning (no device needed): workspaces grow with the problem and hold the transposed planes of the
    tensor-core K1 backward (2 x NG x (Bpad + Cpad) floats on top of the forward's)."""
    fwd = L.exvae_prior_lse_fwd_workspace_bytes(512, 25000, 40)
    both = L.exvae_prior_lse_workspace_bytes(512, 25000, 40)
    assert both - fwd >= 2 * 48 * (512 + 25088) * 4
    assert L.exvae_dense_fwd_workspace_bytes(25512, 784, 300, 1) >= 600 * 784 * 4     # [Wh ; Wg] as one operand
    assert L.exvae_dense_fwd_workspace_bytes(25512, 300, 40, 0) == 0                   # linear layers read x and W in place
    small = L.exvae_gated_dense_bwd_workspace_bytes(512, 300, 300)
    big = L.exvae_gated_dense_bwd_workspace_bytes(25512, 300, 300)
    assert big > small > 512 * 600 * 4


def test_vampprior_state_dict_keys_match_the_reference(golden):
    """prior == 'vampprior': same parameter names and shapes as the reference model that produced the golden
    (means.linear.weight [P, C] for the pseudo-inputs; idle_input is not part of the state_dict there either)."""
    import exemplar_vae_b200 as E
    from oracle import exvae_oracle as O
    g = golden("vamp_step")
    side = int(g["side"])
    args = O.make_args(model_name="vae", prior="vampprior", hidden_size=int(g["hidden"]), number_components=len(g["ex_idx"]),
                       training_set_size=int(g["T"]), input_size=[1, side, side])
    model = E.importing_model(args)(args)
    want = {k[2:]: tuple(v.shape) for k, v in g.items() if k.startswith("p:")}
    got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert got == want
    assert "prior_log_variance" not in got and got["means.linear.weight"] == (side * side, len(g["ex_idx"]))


def test_gated_dense_keeps_its_two_weights_adjacent():
    """layers.GatedDense stores h.weight / g.weight in one [2*O, K] buffer (the GEMM reads [Wh ; Wg] as one operand, no
    concat launch); names, shapes, state_dict and load_state_dict are the reference's (utils/nn.py:44-69), and a module
    that lost the adjacency (deepcopy) is still a valid module (the C side copies then)."""
    import copy
    from exemplar_vae_b200.layers import GatedDense
    m = GatedDense(784, 300)
    adj = lambda mod: mod.g.weight.data_ptr() == mod.h.weight.data_ptr() + mod.h.weight.numel() * 4
    assert adj(m) and m.h.weight.is_contiguous() and m.g.weight.is_contiguous()
    assert list(m.state_dict().keys()) == ["h.weight", "h.bias", "g.weight", "g.bias"]
    sd = {k: torch.randn_like(v) for k, v in m.state_dict().items()}
    m.load_state_dict(sd)
    assert adj(m) and torch.equal(m.h.weight, sd["h.weight"]) and torch.equal(m.g.weight, sd["g.weight"])
    m2 = m.double().float()                      # Module._apply moves every parameter on its own: re-packed afterwards
    assert adj(m2) and torch.equal(m2.g.weight, sd["g.weight"])
    m3 = copy.deepcopy(m)
    assert torch.equal(m3.h.weight, m.h.weight) and torch.equal(m3.g.weight, m.g.weight)
    opt = torch.optim.SGD(m.parameters(), lr=0.5)
    (m.h.weight.sum() + 2 * m.g.weight.sum()).backward()
    opt.step()                                   # in-place updates through the two views hit the shared buffer
    assert torch.allclose(m.h.weight, sd["h.weight"] - 0.5) and torch.allclose(m.g.weight, sd["g.weight"] - 1.0) and adj(m)


def test_grad_projections_see_every_element_and_orderings():
    """oracle.grad_projections (the whole-tensor check of the compact conv goldens): deterministic per name, linear,
    and a transposed filter / swapped channel moves it by O(||g||)."""
    from oracle import exvae_oracle as O
    rs = np.random.RandomState(0)
    g = rs.randn(64, 32, 3, 3)
    p = O.grad_projections(g, "q_z_layers.0.h.weight")
    assert np.array_equal(p, O.grad_projections(g.copy(), "q_z_layers.0.h.weight"))
    assert not np.allclose(p, O.grad_projections(g, "q_z_layers.0.g.weight"))
    assert np.allclose(O.grad_projections(2 * g, "q_z_layers.0.h.weight"), 2 * p)
    nrm = np.linalg.norm(g)
    for wrong in (g.transpose(0, 1, 3, 2), g[:, ::-1], g[::-1]):
        assert np.abs(O.grad_projections(np.ascontiguousarray(wrong), "q_z_layers.0.h.weight") - p).max() > 0.05 * nrm
    e = np.zeros_like(g); e[63, 31, 2, 2] = 1.0             # the very last element enters every projection
    assert np.all(np.abs(O.grad_projections(e, "x")) == 1.0)
