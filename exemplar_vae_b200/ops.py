"""torch.autograd bindings of the exvae_b200 C ABI.

PyTorch is plumbing here: it owns device memory, the current stream and the autograd tape;
every arithmetic step is a kernel of ``csrc/libexvae_b200.so`` launched on
``torch.cuda.current_stream()`` (so whole training steps can be captured in a CUDA graph).
Inputs must be CUDA tensors — a CPU tensor raises; there is no CPU code path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ._lib import ACT_HARDTANH, ACT_NONE, ACT_RELU, ACT_SIGMOID, ExvaeError, lib

__all__ = [
    "prior_lse", "vamp_lse", "vamp_logprob_matrix", "pairwise_distance", "log_normal_diag_vectorized", "prior_logprob_matrix", "knn_topk", "knn_merge",
    "unique_positions", "gather_rows", "scatter_rows_", "gated_dense", "linear", "reparameterize",
    "log_normal_diag", "log_normal_standard", "log_bernoulli", "log_logistic_256", "elbo_reduce",
    "rng_bernoulli", "rng_normal", "rng_randint", "launch_count", "reset_launch_count",
]

_LAUNCHES = 0  # number of exvae kernels enqueued through this module (bench.py reports it)


def gemm_backend() -> str:
    """'tcgen05-3xtf32' when the tensor-core GEMM is active on the current device, else 'fp32-fma'."""
    return "tcgen05-3xtf32" if lib().exvae_gemm_backend() == 1 else "fp32-fma"


def launch_count() -> int:
    return _LAUNCHES


def reset_launch_count() -> None:
    global _LAUNCHES
    _LAUNCHES = 0


def _count(n: int) -> None:
    global _LAUNCHES
    _LAUNCHES += n


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32(t: torch.Tensor, name: str = "tensor") -> torch.Tensor:
    if not t.is_cuda:
        raise ExvaeError(f"{name} must be a CUDA tensor: exemplar_vae_b200 has no CPU path")
    if t.dtype != torch.float32:
        raise ExvaeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def _i64(t: Optional[torch.Tensor], name: str = "index") -> Optional[torch.Tensor]:
    if t is None:
        return None
    if not t.is_cuda:
        raise ExvaeError(f"{name} must be a CUDA tensor")
    if t.dtype != torch.int64:
        t = t.long()
    return t.contiguous()


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ======================================================================================
# K1: fused exemplar prior
# ======================================================================================
_prior_fwd_event = None


def _mark_prior_fwd_done():
    """Event behind the K1 forward when the backward follows in the same call (``g_known``): a consumer of log p(z)
    on another stream waits for THIS event (``take_prior_fwd_event``), not for the whole stream."""
    global _prior_fwd_event
    _prior_fwd_event = torch.cuda.Event()
    _prior_fwd_event.record(torch.cuda.current_stream())


def take_prior_fwd_event():
    global _prior_fwd_event
    ev, _prior_fwd_event = _prior_fwd_event, None
    return ev


class _PriorLSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, mu, logvar, z_idx, mu_idx, c_total, group, c_valid, g_known=None):
        L = lib()
        z, mu, logvar = _f32(z, "z"), _f32(mu, "mu"), _f32(logvar, "logvar")
        B, D = z.shape
        C = mu.shape[0]
        assert mu.shape[1] == D and logvar.numel() == D
        if c_valid is not None:
            assert c_valid.is_cuda and c_valid.dtype == torch.int32 and c_valid.numel() == 1 and group is None
        z_idx = _i64(z_idx)
        mu_idx = _i64(mu_idx)
        if z_idx is not None:
            z_idx = z_idx.reshape(-1)
            assert z_idx.numel() == B
        if mu_idx is not None:
            mu_idx = mu_idx.reshape(-1)
            assert mu_idx.numel() == C
        need_bwd = any(ctx.needs_input_grad)
        ws = _ws(L.exvae_prior_lse_workspace_bytes(B, C, D) if need_bwd else
                 L.exvae_prior_lse_fwd_workspace_bytes(B, C, D), z.device)
        stats = torch.empty((B, 4), dtype=torch.float32, device=z.device)
        log_p = torch.empty((B,), dtype=torch.float32, device=z.device)
        lse2 = torch.empty((B,), dtype=torch.float32, device=z.device)
        total = int(c_total) if c_total is not None else C
        masked = z_idx is not None and mu_idx is not None
        if group is None:
            # one GPU: fwd returns log p(z) itself -- for D <= 63 ONE kernel from the raw inputs (csrc/prior_fused.cu)
            L.check(L.exvae_prior_lse_fwd(_p(z), _p(mu), _p(logvar), _p(z_idx), _p(mu_idx), B, C, D, _p(c_valid), _p(stats),
                                          total, _p(log_p), _p(lse2), _p(ws), ws.numel(), _stream()), "prior_lse_fwd")
            _count(1 if L.exvae_prior_lse_fwd_prepares_ws(B, C, D) == 0 else (5 if masked else 4))
        else:
            import torch.distributed as dist
            L.check(L.exvae_prior_lse_fwd(_p(z), _p(mu), _p(logvar), _p(z_idx), _p(mu_idx), B, C, D, _p(c_valid), _p(stats),
                                          0, None, None, _p(ws), ws.numel(), _stream()), "prior_lse_fwd")
            G = dist.get_world_size(group)
            all_stats = torch.empty((G, B, 4), dtype=torch.float32, device=z.device)
            dist.all_gather_into_tensor(all_stats, stats, group=group)   # the single LSE-partial exchange
            L.check(L.exvae_prior_lse_finalize(_p(all_stats), G, _p(z), _p(logvar), B, D, total, _p(c_valid), _p(log_p),
                                               _p(lse2), _stream()), "prior_lse_finalize")
            _count(3)
        ctx.ws_prepared = int(L.exvae_prior_lse_fwd_prepares_ws(B, C, D))
        ctx.dims = (B, C, D)
        ctx.eager = None
        if g_known is not None and need_bwd and group is None:
            # the upstream gradient of every row is known before the loss exists (a training step whose loss is
            # mean(-RE + beta*KL): d loss / d log p(z_b) = -beta/B): run the backward HERE, on the forward's stream,
            # so that it overlaps the decoder instead of sitting between the loss and the encoder backward
            g = _f32(g_known, "g_known").reshape(-1)
            assert g.numel() == B
            _mark_prior_fwd_done()
            ctx.eager = _PriorLSE._launch_bwd(z, mu, logvar, z_idx, mu_idx, lse2, ws, c_valid, g, (B, C, D), ctx.ws_prepared)
            ctx.lv_shape = logvar.shape
            # the kernels above are only QUEUED: every buffer they read must outlive this call (a [D] logvar row made on
            # the main stream would otherwise be freed, and reused by the decoder, before the prior branch has run)
            ctx.save_for_backward(z, mu, logvar, z_idx, mu_idx, lse2, ws, c_valid, g)
            return log_p
        ctx.save_for_backward(z, mu, logvar, z_idx, mu_idx, lse2, ws, c_valid)
        return log_p

    @staticmethod
    def _launch_bwd(z, mu, logvar, z_idx, mu_idx, lse2, ws, c_valid, g, dims, ws_prepared):
        L = lib()
        B, C, D = dims
        dz = torch.empty_like(z)
        dmu = torch.empty_like(mu)
        dlv = torch.empty((D,), dtype=torch.float32, device=z.device)
        L.check(L.exvae_prior_lse_bwd(_p(z), _p(mu), _p(logvar), _p(z_idx), _p(mu_idx), B, C, D, _p(lse2), _p(g),
                                      _p(dz), _p(dmu), _p(dlv), _p(ws), ws.numel(), ws_prepared, _p(c_valid), _stream()),
                "prior_lse_bwd")
        # D <= 63: prep, pass 1, pass 2 (one launch per wave of bank tiles), rows, dlogvar; D >= 64: W pass, two GEMMs, rows,
        # cols, dlogvar
        extra = 0
        if D <= 63:
            sms = torch.cuda.get_device_properties(z.device).multi_processor_count
            extra = max(0, -(-(-(-C // 128)) // sms) - 1)
        _count((5 if D <= 63 else 6) + extra + (0 if ws_prepared else 1))
        return dz, dmu, dlv

    @staticmethod
    def backward(ctx, g):
        if ctx.eager is not None:
            dz, dmu, dlv = ctx.eager          # computed in forward() from the known upstream gradient
            ctx.eager = None
            return dz, dmu, dlv.view(ctx.lv_shape), None, None, None, None, None, None
        z, mu, logvar, z_idx, mu_idx, lse2, ws, c_valid = ctx.saved_tensors
        g = _f32(g, "grad")
        dz, dmu, dlv = _PriorLSE._launch_bwd(z, mu, logvar, z_idx, mu_idx, lse2, ws, c_valid, g, ctx.dims, ctx.ws_prepared)
        return dz, dmu, dlv.view_as(logvar), None, None, None, None, None, None


def prior_lse(z, mu, logvar, z_idx=None, mu_idx=None, c_total=None, group=None, c_valid=None, g_known=None) -> torch.Tensor:
    """log p(z_b) = LSE_n log N(z_b | mu_n, exp(logvar)) - log(C - #masked_b)   -> [B].

    ``logvar`` is the [D] log-variance vector shared by all exemplars.  ``z_idx``/``mu_idx``
    enable the leave-one-out mask.  With ``group`` the bank ``mu`` is this rank's shard of a
    range-sharded bank: partial (max, sum, count) statistics are all-gathered once and merged;
    ``c_total`` is the global exemplar count.  dz/dlogvar gradients are then per-shard partials.
    ``c_valid`` ([1] int32 on the device): only the first ``c_valid`` bank rows count (fixed-capacity bank of the
    kNN mode); the normaliser then uses that count.
    ``g_known`` ([B] on the device): the gradient of the loss w.r.t. every log p(z_b), when the caller knows it before
    the loss is evaluated (fused training step); the backward kernels then run inside the forward call and
    ``backward`` only hands their results out (its incoming gradient is NOT looked at)."""
    return _PriorLSE.apply(z, mu, logvar, z_idx, mu_idx, c_total, group, c_valid, g_known)


# ======================================================================================
# materialising primitives + kNN (no autograd: used for selection only)
# ======================================================================================
@torch.no_grad()
def pairwise_distance(z, means) -> torch.Tensor:
    L = lib()
    z, means = _f32(z), _f32(means)
    B, D = z.shape
    C = means.shape[0]
    out = torch.empty((B, C), dtype=torch.float32, device=z.device)
    L.check(L.exvae_pairwise_distance(_p(z), _p(means), B, C, D, _p(out), _stream()), "pairwise_distance")
    _count(1)
    return out


@torch.no_grad()
def log_normal_diag_vectorized(x, mean, log_var) -> Tuple[torch.Tensor, torch.Tensor]:
    L = lib()
    x, mean = _f32(x), _f32(mean)
    lv = _f32(log_var).reshape(-1)
    B, D = x.shape
    C = mean.shape[0]
    assert lv.numel() == D
    ln = torch.empty((B, C), dtype=torch.float32, device=x.device)
    pd = torch.empty((B, C), dtype=torch.float32, device=x.device)
    L.check(L.exvae_log_normal_diag_vectorized(_p(x), _p(mean), _p(lv), B, C, D, _p(ln), _p(pd), _stream()),
            "log_normal_diag_vectorized")
    _count(1)
    return ln, pd


@torch.no_grad()
def prior_logprob_matrix(z, mu, logvar, z_idx=None, mu_idx=None) -> torch.Tensor:
    L = lib()
    z, mu, logvar = _f32(z), _f32(mu), _f32(logvar).reshape(-1)
    B, D = z.shape
    C = mu.shape[0]
    z_idx = _i64(z_idx)
    mu_idx = _i64(mu_idx)
    if z_idx is not None:
        z_idx = z_idx.reshape(-1)
    if mu_idx is not None:
        mu_idx = mu_idx.reshape(-1)
    out = torch.empty((B, C), dtype=torch.float32, device=z.device)
    counts = torch.empty((B,), dtype=torch.int32, device=z.device)
    L.check(L.exvae_prior_logprob_matrix(_p(z), _p(mu), _p(logvar), _p(z_idx), _p(mu_idx), B, C, D, _p(out),
                                         _p(counts), _stream()), "prior_logprob_matrix")
    _count(2 if (z_idx is not None and mu_idx is not None) else 1)
    return out


# ======================================================================================
# VampPrior (per-component mean and log-variance)
# ======================================================================================
@torch.no_grad()
def vamp_logprob_matrix(z, mean, logvar) -> torch.Tensor:
    """models/BaseModel.py:84-96 — the [B,C] matrix log N(z_b | mean_c, exp(logvar_c)) - log C."""
    L = lib()
    z, mean, logvar = _f32(z, "z"), _f32(mean, "mean"), _f32(logvar, "logvar")
    B, D = z.shape
    C = mean.shape[0]
    out = torch.empty((B, C), dtype=torch.float32, device=z.device)
    L.check(L.exvae_vamp_logprob_matrix(_p(z), _p(mean), _p(logvar), B, C, D, _p(out), _stream()), "vamp_logprob_matrix")
    _count(1)
    return out


class _VampLSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, mean, logvar):
        L = lib()
        z, mean, logvar = _f32(z, "z"), _f32(mean, "mean"), _f32(logvar, "logvar")
        B, D = z.shape
        C = mean.shape[0]
        mat = torch.empty((B, C), dtype=torch.float32, device=z.device)
        log_p = torch.empty((B,), dtype=torch.float32, device=z.device)
        L.check(L.exvae_vamp_lse_fwd(_p(z), _p(mean), _p(logvar), B, C, D, _p(mat), _p(log_p), _stream()), "vamp_lse_fwd")
        _count(2)
        ctx.save_for_backward(z, mean, logvar, mat, log_p)
        return log_p

    @staticmethod
    def backward(ctx, g):
        L = lib()
        z, mean, logvar, mat, log_p = ctx.saved_tensors
        g = _f32(g)
        B, D = z.shape
        C = mean.shape[0]
        dz, dmean, dlogvar = torch.empty_like(z), torch.empty_like(mean), torch.empty_like(logvar)
        L.check(L.exvae_vamp_lse_bwd(_p(z), _p(mean), _p(logvar), _p(mat), _p(log_p), _p(g), B, C, D, _p(dz), _p(dmean),
                                     _p(dlogvar), _stream()), "vamp_lse_bwd")
        _count(2)
        return dz, dmean, dlogvar


def vamp_lse(z, mean, logvar) -> torch.Tensor:
    """log p(z) under the VampPrior: log-sum-exp over the C components (BaseModel.py:111-128, sum=True)."""
    return _VampLSE.apply(z, mean, logvar)


@torch.no_grad()
def knn_topk(z, bank, k: int, metric: int = 0, pos_offset: int = 0):
    """k smallest distances per row -> (positions [B,k] int64, distances [B,k]).  metric 0 = the
    reference's fp64-expansion squared distance (BaseModel.py:263), 1 = sqrt-Euclidean
    (knn_on_latent.py:4-9)."""
    L = lib()
    z, bank = _f32(z), _f32(bank)
    B, D = z.shape
    C = bank.shape[0]
    idx = torch.empty((B, k), dtype=torch.int64, device=z.device)
    dist = torch.empty((B, k), dtype=torch.float32, device=z.device)
    ws = _ws(L.exvae_knn_workspace_bytes(B, C, D, k), z.device)
    L.check(L.exvae_knn_topk(_p(z), _p(bank), B, C, D, k, metric, pos_offset, _p(idx), _p(dist), _p(ws), ws.numel(),
                             _stream()), "knn_topk")
    _count(1)
    return idx, dist


@torch.no_grad()
def knn_merge(idx, dist):
    """Merge per-shard candidate lists [G,B,k] into the global k smallest."""
    L = lib()
    idx, dist = _i64(idx), _f32(dist)
    G, B, k = idx.shape
    oi = torch.empty((B, k), dtype=torch.int64, device=idx.device)
    od = torch.empty((B, k), dtype=torch.float32, device=idx.device)
    L.check(L.exvae_knn_merge(_p(idx), _p(dist), G, B, k, _p(oi), _p(od), _stream()), "knn_merge")
    _count(1)
    return oi, od


@torch.no_grad()
def unique_positions(pos, value_range: int):
    """Sorted unique values of ``pos`` (all in [0, value_range)) -> (padded [n] int64, count [1] int32 on device)."""
    L = lib()
    pos = _i64(pos).reshape(-1)
    n = pos.numel()
    out = torch.empty((n,), dtype=torch.int64, device=pos.device)
    count = torch.empty((1,), dtype=torch.int32, device=pos.device)
    flags = torch.empty((value_range,), dtype=torch.int32, device=pos.device)
    L.check(L.exvae_unique_positions(_p(pos), n, value_range, _p(out), _p(count), _p(flags), _stream()),
            "unique_positions")
    _count(1)
    return out, count


@torch.no_grad()
def gather_rows(src, idx, out=None) -> torch.Tensor:
    """src[idx] (rows).  ``out`` may be a contiguous [n, row] slice of a larger buffer."""
    L = lib()
    src = _f32(src)
    idx = _i64(idx).reshape(-1)
    n, row = idx.numel(), src.shape[1]
    if out is None:
        out = torch.empty((n, row), dtype=torch.float32, device=src.device)
    assert out.is_contiguous() and tuple(out.shape) == (n, row) and out.dtype == torch.float32
    if n:
        L.check(L.exvae_gather_rows(_p(src), _p(idx), n, row, _p(out), _stream()), "gather_rows")
        _count(1)
    return out


@torch.no_grad()
def gather_index(src, idx) -> torch.Tensor:
    """src[idx] for int64 vectors (exemplars_indices[nearest], models/BaseModel.py:266)."""
    L = lib()
    src, idx = _i64(src).reshape(-1), _i64(idx).reshape(-1)
    out = torch.empty_like(idx)
    if idx.numel():
        L.check(L.exvae_gather_index(_p(src), _p(idx), idx.numel(), _p(out), _stream()), "gather_index")
        _count(1)
    return out


@torch.no_grad()
def scatter_rows_(dst, idx, src) -> torch.Tensor:
    """dst[idx] = src, in place (cache refresh, BaseModel.py:261,269)."""
    L = lib()
    assert dst.is_cuda and dst.dtype == torch.float32 and dst.is_contiguous()
    src = _f32(src)
    idx = _i64(idx).reshape(-1)
    if idx.numel():
        L.check(L.exvae_scatter_rows(_p(dst), _p(idx), idx.numel(), dst.shape[1], _p(src), _stream()), "scatter_rows")
        _count(1)
    return dst


# ======================================================================================
# row-block views whose backward does not zero-fill and add full-size tensors
# ======================================================================================
class _SplitRows(torch.autograd.Function):
    """(t[:B], t[B:]); the backward writes both gradients into ONE buffer (autograd's slice backward would
    zero-fill two full-size tensors and add them)."""

    @staticmethod
    def forward(ctx, t, B):
        ctx.B = B
        ctx.shape = t.shape
        return t[:B], t[B:]

    @staticmethod
    def backward(ctx, g0, g1):
        B = ctx.B
        if g0 is None and g1 is None:
            return None, None
        ref = g0 if g0 is not None else g1
        out = torch.empty(ctx.shape, dtype=ref.dtype, device=ref.device)
        if g0 is not None:
            out[:B].copy_(g0)
        else:
            out[:B].zero_()
        if g1 is not None:
            out[B:].copy_(g1)
        else:
            out[B:].zero_()
        return out, None


def split_rows(t: torch.Tensor, B: int):
    return _SplitRows.apply(t, B)


class _SharedRows(torch.autograd.Function):
    """(t, t[:B]) for a tensor consumed once in full and once by its first B rows: the backward adds the small
    gradient into the first rows of the full one in place (that buffer was produced for this node alone by the
    dense layer's backward) instead of materialising a zero-padded copy."""

    @staticmethod
    def forward(ctx, t, B):
        ctx.B = B
        return t.view_as(t), t[:B]

    @staticmethod
    def backward(ctx, g_full, g_head):
        if g_full is None:
            if g_head is None:
                return None, None
            raise ExvaeError("shared_rows: the full-size consumer produced no gradient")
        if g_head is not None:
            if not g_full.is_contiguous():
                g_full = g_full.contiguous()
            head = g_full[:ctx.B]
            if head.is_cuda:
                L = lib()
                g_head = _f32(g_head)
                L.check(L.exvae_lincomb4(_p(head), _p(g_head), None, None, 1.0, 1.0, 0.0, 0.0, head.numel(), _p(head),
                                         _stream()), "shared_rows_bwd")
                _count(1)
            else:                       # view plumbing exercised on CPU by tests/test_abi.py (no kernel involved)
                head.add_(g_head)
        return g_full, None


def shared_rows(t: torch.Tensor, B: int):
    return _SharedRows.apply(t, B)


# ======================================================================================
# K3: dense layers
# ======================================================================================
class _GatedDense(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, Wh, bh, Wg, bg, sink):
        L = lib()
        ctx.sink = sink
        x, Wh, Wg = _f32(x, "x"), _f32(Wh), _f32(Wg)
        bh = _f32(bh) if bh is not None else None
        bg = _f32(bg) if bg is not None else None
        R, K = x.shape
        O = Wh.shape[0]
        out = torch.empty((R, O), dtype=torch.float32, device=x.device)
        need = any(ctx.needs_input_grad)
        s = torch.empty_like(out) if need else None
        # forward workspace = [Wh ; Wg] as one GEMM operand of the tensor-core backend; kept for the backward
        fws = _ws(L.exvae_dense_fwd_workspace_bytes(R, K, O, 1), x.device)
        L.check(L.exvae_gated_dense_fwd(_p(x), _p(Wh), _p(bh), _p(Wg), _p(bg), R, K, O, _p(out), _p(s),
                                        _p(fws), fws.numel(), _stream()), "gated_dense_fwd")
        _count(1 if Wg.data_ptr() == Wh.data_ptr() + 4 * Wh.numel() else 2)     # adjacent weights: no [Wh ; Wg] copy
        ctx.save_for_backward(x, Wh, Wg, out if need else None, s, fws if need else None)
        ctx.has_bias = (bh is not None, bg is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        L = lib()
        x, Wh, Wg, out, s, fws = ctx.saved_tensors
        dout = _f32(dout)
        R, K = x.shape
        O = Wh.shape[0]
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        sink = ctx.sink
        if sink is not None:       # fused accumulation straight into the parameters' .grad storage
            dWh, dbh, dWg, dbg = sink
        else:
            dWh, dWg = torch.empty_like(Wh), torch.empty_like(Wg)
            dbh = torch.empty((O,), dtype=torch.float32, device=x.device) if ctx.has_bias[0] else None
            dbg = torch.empty((O,), dtype=torch.float32, device=x.device) if ctx.has_bias[1] else None
        ws = _ws(L.exvae_gated_dense_bwd_workspace_bytes(R, K, O), x.device)
        L.check(L.exvae_gated_dense_bwd(_p(x), _p(Wh), _p(Wg), _p(out), _p(s), _p(dout), R, K, O, _p(dx), _p(dWh),
                                        _p(dbh), _p(dWg), _p(dbg), _p(fws), fws.numel() if fws is not None else 0,
                                        _p(ws), ws.numel(), 1 if sink is not None else 0, _stream()),
                "gated_dense_bwd")
        _count(3 + (1 if dx is not None else 0))
        if sink is not None:
            _keep_until_flush(ws)
            _sink_done(sink)
            return dx, None, None, None, None, None
        return dx, dWh, dbh, dWg, dbg, None


def _grad_sink(*params):
    """The .grad buffers of ``params`` when fused gradient accumulation is possible (every parameter
    already owns a contiguous .grad, e.g. views of distributed.FlatGrads), else None."""
    if not torch.is_grad_enabled() or not _FUSE_GRAD_ACCUM:
        return None
    out = []
    for p in params:
        if p is None:
            out.append(None)
            continue
        if not p.is_leaf:
            return None
        g = p.grad
        if g is None or not p.requires_grad or not g.is_contiguous() or g.dtype != torch.float32:
            return None
        out.append(g)
    return tuple(out)


_FUSE_GRAD_ACCUM = False
_GRAD_READY_HOOK = None      # called with each .grad buffer a fused dense backward has just completed (distributed.GradBuckets)


def set_grad_ready_hook(fn) -> None:
    global _GRAD_READY_HOOK
    _GRAD_READY_HOOK = fn


_DEFER_FINISH = False
_DEFERRED_WS = []           # workspaces a queued (deferred) dW finish still reads


def set_deferred_dw_finish(on: bool) -> bool:
    """When on (together with fused gradient accumulation), the backward of a LARGE dense layer queues its last pass
    (split-K reduction of dW + bias sums into ``.grad``) on the library's side stream without joining it, so that it runs
    next to the following layer's staging kernel; ``flush_dense_bwd()`` joins.  The caller must flush before anything
    reads a parameter gradient (optimizer step, gradient all-reduce).  Returns the previous setting."""
    global _DEFER_FINISH
    prev = _DEFER_FINISH
    if prev and not on:
        flush_dense_bwd()
    _DEFER_FINISH = bool(on)
    lib().exvae_dense_bwd_defer_finish(1 if on else 0)
    return prev


def flush_dense_bwd() -> None:
    """The current stream waits for every queued dW finish (no-op when none is pending)."""
    lib().check(lib().exvae_dense_bwd_flush(_stream()), "dense_bwd_flush")
    _DEFERRED_WS.clear()


def _keep_until_flush(ws) -> None:
    if _DEFER_FINISH:
        _DEFERRED_WS.append(ws)


def _sink_done(sink) -> None:
    if _GRAD_READY_HOOK is not None:
        if _DEFER_FINISH:
            flush_dense_bwd()          # the hook hands the gradient to the all-reduce: it must be complete
        for g in sink:
            if g is not None:
                _GRAD_READY_HOOK(g, True)


def set_fused_grad_accumulation(on: bool) -> bool:
    """When on, dense layers ADD their weight/bias gradients directly into existing ``.grad`` buffers
    inside the backward kernels (autograd then sees no gradient for those parameters).  Returns the
    previous setting so callers can scope it."""
    global _FUSE_GRAD_ACCUM
    prev = _FUSE_GRAD_ACCUM
    _FUSE_GRAD_ACCUM = bool(on)
    return prev


def gated_dense(x, Wh, bh, Wg, bg) -> torch.Tensor:
    """(x Wh^T + bh) * sigmoid(x Wg^T + bg)   (utils/nn.py:44-69)."""
    return _GatedDense.apply(x, Wh, bh, Wg, bg, _grad_sink(Wh, bh, Wg, bg))


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b, act, lo, hi, sink):
        L = lib()
        ctx.sink = sink
        x, W = _f32(x, "x"), _f32(W)
        b = _f32(b) if b is not None else None
        R, K = x.shape
        O = W.shape[0]
        out = torch.empty((R, O), dtype=torch.float32, device=x.device)
        fws = _ws(L.exvae_dense_fwd_workspace_bytes(R, K, O, 0), x.device)
        L.check(L.exvae_linear_fwd(_p(x), _p(W), _p(b), R, K, O, act, lo, hi, _p(out), _p(fws), fws.numel(),
                                   _stream()), "linear_fwd")
        _count(1)
        ctx.save_for_backward(x, W, out if act != ACT_NONE else None, fws if any(ctx.needs_input_grad) else None)
        ctx.cfg = (act, lo, hi, b is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        L = lib()
        x, W, out, fws = ctx.saved_tensors
        act, lo, hi, has_b = ctx.cfg
        dout = _f32(dout)
        R, K = x.shape
        O = W.shape[0]
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        sink = ctx.sink
        if sink is not None:
            dW, db = sink
        else:
            dW = torch.empty_like(W)
            db = torch.empty((O,), dtype=torch.float32, device=x.device) if has_b else None
        ws = _ws(L.exvae_linear_bwd_workspace_bytes(R, K, O), x.device)
        L.check(L.exvae_linear_bwd(_p(x), _p(W), _p(out), _p(dout), R, K, O, act, lo, hi, _p(dx), _p(dW), _p(db),
                                   _p(fws), fws.numel() if fws is not None else 0, _p(ws), ws.numel(),
                                   1 if sink is not None else 0, _stream()), "linear_bwd")
        _count(3 + (1 if dx is not None else 0))
        if sink is not None:
            _keep_until_flush(ws)
            _sink_done(sink)
            return dx, None, None, None, None, None, None
        return dx, dW, db, None, None, None, None


def linear(x, W, b=None, act: int = ACT_NONE, lo: float = 0.0, hi: float = 0.0) -> torch.Tensor:
    """act(x W^T + b)   (utils/nn.py:29-41)."""
    return _Linear.apply(x, W, b, act, float(lo), float(hi), _grad_sink(W, b))


# ======================================================================================
# element-wise pieces of the ELBO
# ======================================================================================
class _Reparam(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, logvar, eps):
        L = lib()
        mu, logvar, eps = _f32(mu), _f32(logvar), _f32(eps)
        z = torch.empty_like(mu)
        L.check(L.exvae_reparameterize_fwd(_p(mu), _p(logvar), _p(eps), mu.numel(), _p(z), _stream()), "reparam_fwd")
        _count(1)
        ctx.save_for_backward(logvar, eps)
        return z

    @staticmethod
    def backward(ctx, dz):
        L = lib()
        logvar, eps = ctx.saved_tensors
        dz = _f32(dz)
        dmu = torch.empty_like(dz) if ctx.needs_input_grad[0] else None
        dlv = torch.empty_like(dz) if ctx.needs_input_grad[1] else None
        L.check(L.exvae_reparameterize_bwd(_p(logvar), _p(eps), _p(dz), dz.numel(), _p(dmu), _p(dlv), _stream()),
                "reparam_bwd")
        _count(1)
        return dmu, dlv, None


def reparameterize(mu, logvar, eps) -> torch.Tensor:
    return _Reparam.apply(mu, logvar, eps)


class _ReparamLogQ(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, logvar, eps):
        L = lib()
        mu, logvar, eps = _f32(mu), _f32(logvar), _f32(eps)
        B, D = mu.shape
        z = torch.empty_like(mu)
        logq = torch.empty((B,), dtype=torch.float32, device=mu.device)
        L.check(L.exvae_reparam_logq_fwd(_p(mu), _p(logvar), _p(eps), B, D, _p(z), _p(logq), _stream()), "reparam_logq")
        _count(1)
        ctx.save_for_backward(mu, logvar, eps, z)
        return z, logq

    @staticmethod
    def backward(ctx, dz, dlogq):
        L = lib()
        mu, logvar, eps, z = ctx.saved_tensors
        B, D = mu.shape
        dz = _f32(dz) if dz is not None else None
        dlogq = _f32(dlogq) if dlogq is not None else None
        dmu = torch.empty_like(mu) if ctx.needs_input_grad[0] else None
        dlv = torch.empty_like(mu) if ctx.needs_input_grad[1] else None
        if dmu is None and dlv is None:
            return None, None, None
        L.check(L.exvae_reparam_logq_bwd(_p(mu), _p(logvar), _p(eps), _p(z), _p(dz), _p(dlogq), B, D, _p(dmu), _p(dlv),
                                         _stream()), "reparam_logq_bwd")
        _count(1)
        return dmu, dlv, None


def reparam_logq(mu, logvar, eps):
    """(z, log q(z|x)): reparameterize (models/BaseModel.py:79-82) and log_normal_diag(z, mu, logvar, dim=1)
    (utils/distributions.py:28-33) in one kernel each way; bit-identical to the separate calls."""
    return _ReparamLogQ.apply(mu, logvar, eps)


class _Fanout(torch.autograd.Function):
    """n aliases of one tensor whose backward sums the n gradients in ONE exvae kernel (autograd's own accumulation
    would launch n-1 ATen add kernels)."""

    @staticmethod
    def forward(ctx, t, n):
        return tuple(t.view_as(t) for _ in range(n))

    @staticmethod
    def backward(ctx, *gs):
        L = lib()
        gs = [_f32(g) for g in gs if g is not None]
        if not gs:
            return None, None
        if len(gs) == 1:
            return gs[0], None
        out = torch.empty_like(gs[0])
        ptrs = [_p(g) for g in gs] + [None] * (4 - len(gs))
        cs = [1.0] * len(gs) + [0.0] * (4 - len(gs))
        L.check(L.exvae_lincomb4(*ptrs, *cs, out.numel(), _p(out), _stream()), "fanout_bwd")
        _count(1)
        return out, None


def fanout(t: torch.Tensor, n: int):
    assert 2 <= n <= 4
    return _Fanout.apply(t, n)


class _BcastScalar(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s, n, sink):
        L = lib()
        s = _f32(s, "scalar")
        assert s.numel() == 1
        ctx.sink = sink
        out = torch.empty((n,), dtype=torch.float32, device=s.device)
        L.check(L.exvae_bcast_scalar(_p(s), n, _p(out), _stream()), "bcast_scalar")
        _count(1)
        ctx.shape = s.shape
        return out

    @staticmethod
    def backward(ctx, g):
        L = lib()
        g = _f32(g)
        sink = ctx.sink
        dst = sink[0] if sink is not None else torch.empty(ctx.shape, dtype=torch.float32, device=g.device)
        L.check(L.exvae_sum_to_scalar(_p(g), g.numel(), _p(dst), 1 if sink is not None else 0, _stream()), "sum_to_scalar")
        _count(1)
        if sink is not None:
            if _GRAD_READY_HOOK is not None:
                _GRAD_READY_HOOK(sink[0], may_fire=False)     # may run on the prior side stream: count only
            return None, None, None
        return dst, None, None


def bcast_scalar(s: torch.Tensor, n: int) -> torch.Tensor:
    """A [1] parameter as a contiguous [n] row (prior_log_variance -> bank log-variance row); the gradient is summed
    (and, with fused accumulation, added into the parameter's .grad) by one kernel."""
    return _BcastScalar.apply(s, int(n), _grad_sink(s))


class _ConcatCols(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        L = lib()
        a, b = _f32(a), _f32(b)
        R, Ka = a.shape
        Kb = b.shape[1]
        out = torch.empty((R, Ka + Kb), dtype=torch.float32, device=a.device)
        L.check(L.exvae_concat_cols_fwd(_p(a), _p(b), R, Ka, Kb, _p(out), _stream()), "concat_cols")
        _count(1)
        ctx.dims = (R, Ka, Kb)
        return out

    @staticmethod
    def backward(ctx, g):
        L = lib()
        R, Ka, Kb = ctx.dims
        g = _f32(g)
        da = torch.empty((R, Ka), dtype=torch.float32, device=g.device) if ctx.needs_input_grad[0] else None
        db = torch.empty((R, Kb), dtype=torch.float32, device=g.device) if ctx.needs_input_grad[1] else None
        if da is None and db is None:
            return None, None
        L.check(L.exvae_concat_cols_bwd(_p(g), R, Ka, Kb, _p(da), _p(db), _stream()), "concat_cols_bwd")
        _count(1)
        return da, db


def concat_cols(a, b) -> torch.Tensor:
    """torch.cat((a, b), 1) for two [R, *] matrices (models/AbsHModel.py:55,83)."""
    return _ConcatCols.apply(a, b)


@torch.no_grad()
def zero_(t: torch.Tensor) -> torch.Tensor:
    assert t.is_cuda and t.is_contiguous()
    L = lib()
    L.check(L.exvae_zero(_p(t), t.numel() * t.element_size(), _stream()), "zero")
    return t


class _LogNormalDiag(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mean, logvar):
        L = lib()
        x, mean, logvar = _f32(x), _f32(mean), _f32(logvar)
        B, D = x.shape
        out = torch.empty((B,), dtype=torch.float32, device=x.device)
        L.check(L.exvae_log_normal_diag_fwd(_p(x), _p(mean), _p(logvar), B, D, _p(out), _stream()), "log_normal_diag")
        _count(1)
        ctx.save_for_backward(x, mean, logvar)
        return out

    @staticmethod
    def backward(ctx, g):
        L = lib()
        x, mean, logvar = ctx.saved_tensors
        g = _f32(g)
        B, D = x.shape
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dm = torch.empty_like(x) if ctx.needs_input_grad[1] else None
        dl = torch.empty_like(x) if ctx.needs_input_grad[2] else None
        L.check(L.exvae_log_normal_diag_bwd(_p(x), _p(mean), _p(logvar), _p(g), B, D, _p(dx), _p(dm), _p(dl),
                                            _stream()), "log_normal_diag_bwd")
        _count(1)
        return dx, dm, dl


def log_normal_diag(x, mean, log_var) -> torch.Tensor:
    """utils/distributions.py:28-33 with dim=1 (x, mean, log_var all [B, D])."""
    return _LogNormalDiag.apply(x, mean, log_var)


class _LogNormalStd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        L = lib()
        x = _f32(x)
        B, D = x.shape
        out = torch.empty((B,), dtype=torch.float32, device=x.device)
        L.check(L.exvae_log_normal_standard_fwd(_p(x), B, D, _p(out), _stream()), "log_normal_standard")
        _count(1)
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, g):
        L = lib()
        (x,) = ctx.saved_tensors
        g = _f32(g)
        B, D = x.shape
        dx = torch.empty_like(x)
        L.check(L.exvae_log_normal_standard_bwd(_p(x), _p(g), B, D, _p(dx), _stream()), "log_normal_standard_bwd")
        _count(1)
        return dx


def log_normal_standard(x) -> torch.Tensor:
    return _LogNormalStd.apply(x)


class _LogBernoulli(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mean):
        L = lib()
        x, mean = _f32(x), _f32(mean)
        B, P = mean.shape
        out = torch.empty((B,), dtype=torch.float32, device=x.device)
        L.check(L.exvae_log_bernoulli_fwd(_p(x), _p(mean), B, P, _p(out), _stream()), "log_bernoulli")
        _count(1)
        ctx.save_for_backward(x, mean)
        return out

    @staticmethod
    def backward(ctx, g):
        L = lib()
        x, mean = ctx.saved_tensors
        g = _f32(g)
        B, P = mean.shape
        dm = torch.empty_like(mean)
        L.check(L.exvae_log_bernoulli_bwd(_p(x), _p(mean), _p(g), B, P, _p(dm), _stream()), "log_bernoulli_bwd")
        _count(1)
        return None, dm


def log_bernoulli(x, mean) -> torch.Tensor:
    """utils/distributions.py:44-51 with dim=1; gradient flows to ``mean`` only (x is data)."""
    return _LogBernoulli.apply(x, mean)


class _LogLogistic256(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mean, logvar):
        L = lib()
        x, mean, logvar = _f32(x), _f32(mean), _f32(logvar)
        B, P = mean.shape
        out = torch.empty((B,), dtype=torch.float32, device=x.device)
        L.check(L.exvae_log_logistic256_fwd(_p(x), _p(mean), _p(logvar), B, P, _p(out), _stream()), "log_logistic_256")
        _count(1)
        ctx.save_for_backward(x, mean, logvar)
        return out

    @staticmethod
    def backward(ctx, g):
        L = lib()
        x, mean, logvar = ctx.saved_tensors
        g = _f32(g)
        B, P = mean.shape
        dm = torch.empty_like(mean) if ctx.needs_input_grad[1] else None
        dl = torch.empty_like(mean) if ctx.needs_input_grad[2] else None
        L.check(L.exvae_log_logistic256_bwd(_p(x), _p(mean), _p(logvar), _p(g), B, P, _p(dm), _p(dl), _stream()),
                "log_logistic_256_bwd")
        _count(1)
        return None, dm, dl


def log_logistic_256(x, mean, logvar) -> torch.Tensor:
    return _LogLogistic256.apply(x, mean, logvar)


class _ElboReduce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, RE, KL, beta, average):
        L = lib()
        RE, KL = _f32(RE), _f32(KL)
        B = RE.numel()
        # beta is a host float, or a [1] fp32 device tensor the kernels read at run time (a captured graph then
        # follows the warm-up schedule of utils/training.py:5-12 without being rebuilt)
        beta_dev = beta if torch.is_tensor(beta) else None
        if beta_dev is not None:
            assert beta_dev.is_cuda and beta_dev.dtype == torch.float32 and beta_dev.numel() == 1
        beta_f = 0.0 if beta_dev is not None else float(beta)
        ctx.cfg = (B, beta_f, bool(average))
        ctx.beta_dev = beta_dev
        if average:
            out3 = torch.empty((3,), dtype=torch.float32, device=RE.device)
            L.check(L.exvae_elbo_reduce(_p(RE), _p(KL), B, beta_f, _p(beta_dev), 1, _p(out3), None, _stream()),
                    "elbo_reduce")
            _count(1)
            return out3
        loss_b = torch.empty((B,), dtype=torch.float32, device=RE.device)
        L.check(L.exvae_elbo_reduce(_p(RE), _p(KL), B, beta_f, _p(beta_dev), 0, None, _p(loss_b), _stream()),
                "elbo_reduce")
        _count(1)
        return loss_b

    @staticmethod
    def backward(ctx, g):
        L = lib()
        B, beta, average = ctx.cfg
        g = _f32(g)
        dRE = torch.empty((B,), dtype=torch.float32, device=g.device)
        dKL = torch.empty((B,), dtype=torch.float32, device=g.device)
        L.check(L.exvae_elbo_reduce_bwd(_p(g) if average else None, None if average else _p(g), B, beta,
                                        _p(ctx.beta_dev), 1 if average else 0, _p(dRE), _p(dKL), _stream()),
                "elbo_reduce_bwd")
        _count(1)
        return dRE, dKL, None, None


def elbo_reduce(RE, KL, beta, average: bool):
    """models/BaseModel.py:71-75.  average=True -> tensor [3] = (mean loss, mean RE, mean KL);
    average=False -> per-sample loss [B].  ``beta``: float or [1] fp32 device tensor."""
    return _ElboReduce.apply(RE, KL, beta if torch.is_tensor(beta) else float(beta), bool(average))


# ======================================================================================
# counter-based RNG
# ======================================================================================
@torch.no_grad()
def _adv(counter, advance) -> int:
    """advance=True lets the drawing kernel bump counter[0] itself (needs the 2-word counter of DeviceRng)."""
    if not advance:
        return 0
    if counter is None or counter.numel() < 2:
        raise ExvaeError("in-kernel RNG advance needs a [2] int64 counter tensor (offset, ticket)")
    return 1


@torch.no_grad()
def rng_bernoulli(p, seed: int, counter: Optional[torch.Tensor], subseq: int, advance: bool = False) -> torch.Tensor:
    L = lib()
    p = _f32(p)
    out = torch.empty_like(p)
    L.check(L.exvae_rng_bernoulli(_p(p), p.numel(), seed, _p(counter), subseq, _adv(counter, advance), _p(out),
                                  _stream()), "rng_bernoulli")
    _count(1)
    return out


@torch.no_grad()
def rng_normal(shape, seed: int, counter: Optional[torch.Tensor], subseq: int, device,
               advance: bool = False) -> torch.Tensor:
    L = lib()
    out = torch.empty(shape, dtype=torch.float32, device=device)
    L.check(L.exvae_rng_normal(out.numel(), seed, _p(counter), subseq, _adv(counter, advance), _p(out), _stream()),
            "rng_normal")
    _count(1)
    return out


@torch.no_grad()
def rng_randint(low: int, high: int, n: int, seed: int, counter: Optional[torch.Tensor], subseq: int, device,
                advance: bool = False):
    L = lib()
    out = torch.empty((n,), dtype=torch.int64, device=device)
    L.check(L.exvae_rng_randint(low, high, n, seed, _p(counter), subseq, _adv(counter, advance), _p(out), _stream()),
            "rng_randint")
    _count(1)
    return out


@torch.no_grad()
def rng_advance_(counter: torch.Tensor, by: int = 1) -> None:
    L = lib()
    L.check(L.exvae_rng_advance(_p(counter), by, _stream()), "rng_advance")
    _count(1)


class _LinComb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coeffs, *xs):
        L = lib()
        xs = [_f32(x) for x in xs]
        assert 1 <= len(xs) <= 4 and len(coeffs) == len(xs)
        ctx.coeffs = tuple(float(c) for c in coeffs)
        out = torch.empty_like(xs[0])
        ptrs = [_p(x) for x in xs] + [None] * (4 - len(xs))
        cs = list(ctx.coeffs) + [0.0] * (4 - len(xs))
        L.check(L.exvae_lincomb4(*ptrs, *cs, out.numel(), _p(out), _stream()), "lincomb4")
        _count(1)
        return out

    @staticmethod
    def backward(ctx, g):
        L = lib()
        g = _f32(g)
        grads = []
        for i, c in enumerate(ctx.coeffs):
            if not ctx.needs_input_grad[i + 1]:
                grads.append(None)
                continue
            if c == 1.0:
                grads.append(g)        # the upstream gradient itself: no kernel
                continue
            d = torch.empty_like(g)
            L.check(L.exvae_lincomb4(_p(g), None, None, None, c, 0.0, 0.0, 0.0, g.numel(), _p(d), _stream()),
                    "lincomb4_bwd")
            _count(1)
            grads.append(d)
        return (None, *grads)


def lincomb(coeffs, *xs) -> torch.Tensor:
    """sum_j coeffs[j] * xs[j] for up to four same-shaped tensors (KL assembly)."""
    return _LinComb.apply(tuple(coeffs), *xs)


# ======================================================================================
# K1 over a range-sharded exemplar bank (one process per GPU)
# ======================================================================================
class _PriorLSESharded(torch.autograd.Function):
    """Every rank holds B local latents and its own shard of the bank.  Forward: all-gather the
    latents (+indices), run K1 for ALL rows against the LOCAL shard, all-gather the [B_total,4]
    partial statistics (the single LSE exchange) and merge.  Backward: all-gather the row grads,
    run the K1 backward against the local shard (dmu is complete for the shard), reduce-scatter
    the partial dz back to the row owners.

    The exchanges are single multimem kernels over the NVSwitch multicast mapping of a symmetric arena
    (distributed.McComm, csrc/mc_coll.cu) when the box supports it, else NCCL collectives."""

    @staticmethod
    def forward(ctx, z, mu, logvar, z_idx, mu_idx, c_total, group, g_known=None):
        import torch.distributed as dist
        from .distributed import mc_comm
        L = lib()
        z, mu, logvar = _f32(z, "z"), _f32(mu, "mu"), _f32(logvar, "logvar")
        G, rank = dist.get_world_size(group), dist.get_rank(group)
        B, D = z.shape
        C = mu.shape[0]
        masked = z_idx is not None and mu_idx is not None
        comm = mc_comm(group)
        if comm is not None and ((B * D) % 4 or B % 4):
            comm = None                                    # the multimem kernels move 16-byte units
        zi_all = None
        if masked:
            z_idx = _i64(z_idx).reshape(-1)
            mu_idx = _i64(mu_idx).reshape(-1)
        else:
            mu_idx = None
        if comm is not None:
            z_all = comm.all_gather(z, "z").view(G * B, D)
            if masked:
                zi_all = comm.all_gather(z_idx, "zi").view(G * B)
            _count(2 if masked else 1)
        else:
            z_all = torch.empty((G * B, D), dtype=torch.float32, device=z.device)
            # (torch's _coalescing_manager would make the two gathers one NCCL launch, but it breaks CUDA-graph
            #  capture: "dependency created on uncaptured work in another stream", torch 2.11)
            dist.all_gather_into_tensor(z_all, z, group=group)
            if masked:
                zi_all = torch.empty((G * B,), dtype=torch.int64, device=z.device)
                dist.all_gather_into_tensor(zi_all, z_idx, group=group)
        Bt = G * B
        ws = _ws(L.exvae_prior_lse_workspace_bytes(Bt, C, D), z.device)
        stats = torch.empty((Bt, 4), dtype=torch.float32, device=z.device)
        L.check(L.exvae_prior_lse_fwd(_p(z_all), _p(mu), _p(logvar), _p(zi_all), _p(mu_idx), Bt, C, D, None, _p(stats),
                                      0, None, None, _p(ws), ws.numel(), _stream()), "prior_lse_fwd")
        prepared = int(L.exvae_prior_lse_fwd_prepares_ws(Bt, C, D))
        _count(3)
        if comm is not None:
            all_stats = comm.all_gather(stats, "stats")             # the single LSE-partial exchange: one kernel
            _count(1)
        else:
            all_stats = torch.empty((G, Bt, 4), dtype=torch.float32, device=z.device)
            dist.all_gather_into_tensor(all_stats, stats, group=group)
        log_p = torch.empty((Bt,), dtype=torch.float32, device=z.device)
        lse2 = torch.empty((Bt,), dtype=torch.float32, device=z.device)
        L.check(L.exvae_prior_lse_finalize(_p(all_stats), G, _p(z_all), _p(logvar), Bt, D, int(c_total), None, _p(log_p),
                                           _p(lse2), _stream()), "prior_lse_finalize")
        _count(1)
        ctx.meta = (B, Bt, C, D, G, rank, group, comm, prepared)
        ctx.eager = None
        out = log_p[rank * B:(rank + 1) * B].clone()
        if g_known is not None and any(ctx.needs_input_grad):
            # known upstream gradient of ALL G*B rows (see prior_lse): backward + dz reduce-scatter run here, and the
            # all-gather of the row gradients disappears
            g_all = _f32(g_known, "g_known").reshape(-1)
            assert g_all.numel() == Bt
            _mark_prior_fwd_done()
            ctx.eager = _PriorLSESharded._launch_bwd(z_all, mu, logvar, zi_all, mu_idx, lse2, ws, ctx.meta, None, g_all)
            ctx.lv_shape = logvar.shape
            ctx.save_for_backward(z_all, mu, logvar, zi_all, mu_idx, lse2, ws, g_all)     # keep the queued kernels' inputs alive
        else:
            ctx.save_for_backward(z_all, mu, logvar, zi_all, mu_idx, lse2, ws)
        return out

    @staticmethod
    def _launch_bwd(z_all, mu, logvar, zi_all, mu_idx, lse2, ws, meta, g, g_all):
        import torch.distributed as dist
        L = lib()
        B, Bt, C, D, G, rank, group, comm, prepared = meta
        dev = z_all.device
        dmu = torch.empty_like(mu)
        dlv = torch.empty((D,), dtype=torch.float32, device=dev)
        dz = torch.empty((B, D), dtype=torch.float32, device=dev)
        if comm is not None:
            if g_all is None:
                g_all = comm.all_gather(g, "g").view(Bt)
            dz_all, dz_off = comm.region("dz", Bt * D)               # K1 writes its partial dz straight into the arena
            dz_all = dz_all.view(Bt, D)
        else:
            if g_all is None:
                g_all = torch.empty((Bt,), dtype=torch.float32, device=dev)
                dist.all_gather_into_tensor(g_all, g, group=group)
            dz_all = torch.empty_like(z_all)
        L.check(L.exvae_prior_lse_bwd(_p(z_all), _p(mu), _p(logvar), _p(zi_all), _p(mu_idx), Bt, C, D, _p(lse2),
                                      _p(g_all), _p(dz_all), _p(dmu), _p(dlv), _p(ws), ws.numel(), prepared, None, _stream()),
                "prior_lse_bwd")
        _count(3)
        if comm is not None:
            comm.reduce_scatter(dz_off, dz)                          # sum over the shards, reduced in the switch
            _count(2)
        else:
            dist.reduce_scatter_tensor(dz, dz_all, op=dist.ReduceOp.SUM, group=group)
        return dz, dmu, dlv

    @staticmethod
    def backward(ctx, g):
        if ctx.eager is not None:
            dz, dmu, dlv = ctx.eager          # computed in forward() from the known upstream gradient
            ctx.eager = None
            return dz, dmu, dlv.view(ctx.lv_shape), None, None, None, None, None
        z_all, mu, logvar, zi_all, mu_idx, lse2, ws = ctx.saved_tensors
        dz, dmu, dlv = _PriorLSESharded._launch_bwd(z_all, mu, logvar, zi_all, mu_idx, lse2, ws, ctx.meta,
                                                    _f32(g, "grad"), None)
        return dz, dmu, dlv.view_as(logvar), None, None, None, None, None


def prior_lse_sharded(z, mu_shard, logvar, z_idx, mu_idx_shard, c_total: int, group, g_known=None) -> torch.Tensor:
    """``prior_lse`` for a bank range-sharded over ``group``: returns log p(z_b) for the LOCAL rows.
    Gradients: dz local (summed over shards), dmu for the local shard, dlogvar = this shard's share
    (the data-parallel gradient all-reduce completes it, like every other replicated parameter)."""
    return _PriorLSESharded.apply(z, mu_shard, logvar, z_idx, mu_idx_shard, c_total, group, g_known)


# ======================================================================================
# K4: convolution (NHWC): implicit GEMM on tcgen05 (>= 16 input channels), patch-matrix GEMM otherwise; ELU; 2x upsample
# ======================================================================================
def _conv_out(h, k, s, p):
    return (h + 2 * p - k) // s + 1


def _conv_plan(N, H, W, Cin, KH, KW, stride, pad, ncat):
    import ctypes
    L = lib()
    arr = (ctypes.c_int * 9)()
    L.check(L.exvae_conv_plan(N, H, W, Cin, KH, KW, stride, pad, ncat, ctypes.addressof(arr)), "conv_plan")
    keys = ("OH", "OW", "implicit", "cpad", "Kp", "Kpc", "dx_implicit", "cpad_dx", "Kp_dx")
    return dict(zip(keys, list(arr)))


def _pack_filters(W0, W1, mode, cpad, Kp, rows_total):
    """[rows_total][Kp] operand from the reference-layout filters ([Cout,Cin,KH,KW]); gated layers: h then g."""
    L = lib()
    O, Cin, KH, KW = W0.shape
    out = torch.empty((rows_total, Kp), dtype=torch.float32, device=W0.device)
    L.check(L.exvae_conv_pack_weight(_p(W0), O, Cin, KH, KW, mode, cpad, Kp, 0, 1, rows_total, _p(out), _stream()),
            "conv_pack_weight")
    _count(1)
    if W1 is not None:
        L.check(L.exvae_conv_pack_weight(_p(W1), O, Cin, KH, KW, mode, cpad, Kp, O, 0, rows_total, _p(out), _stream()),
                "conv_pack_weight")
        _count(1)
    return out


class _Conv2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W0, b0, W1, b1, stride, pad, act, lo, hi, sink):
        L = lib()
        x, W0 = _f32(x, "x"), _f32(W0, "weight")
        gated = W1 is not None
        W1 = _f32(W1) if gated else None
        b0 = _f32(b0) if b0 is not None else None
        b1 = _f32(b1) if b1 is not None else None
        N, H, Wd, C = x.shape
        O, Cin, KH, KW = W0.shape
        assert Cin == C
        ncat = 2 * O if gated else O
        pl = _conv_plan(N, H, Wd, C, KH, KW, stride, pad, ncat)
        wpk = _pack_filters(W0, W1, 0, pl["cpad"], pl["Kp"], ncat)
        out = torch.empty((N, pl["OH"], pl["OW"], O), dtype=torch.float32, device=x.device)
        need = any(ctx.needs_input_grad)
        sig = torch.empty_like(out) if (gated and need) else None
        ws = _ws(L.exvae_conv2d_fwd_workspace_bytes(N, H, Wd, C, KH, KW, stride, pad, ncat), x.device)
        L.check(L.exvae_conv2d_fwd(_p(x), _p(wpk), _p(b0), _p(b1), N, H, Wd, C, KH, KW, stride, pad, O, 1 if gated else 0,
                                   act, lo, hi, _p(out), _p(sig), _p(ws), ws.numel(), _stream()), "conv2d_fwd")
        _count(1 if pl["implicit"] else 2)
        keep_out = need and (gated or act != ACT_NONE)
        ctx.save_for_backward(x, W0, W1, out if keep_out else None, sig)
        ctx.cfg = (stride, pad, act, lo, hi, b0 is not None, b1 is not None)
        ctx.sink = sink
        return out

    @staticmethod
    def backward(ctx, dout):
        L = lib()
        x, W0, W1, out, sig = ctx.saved_tensors
        stride, pad, act, lo, hi, has_b0, has_b1 = ctx.cfg
        gated = W1 is not None
        dout = _f32(dout)
        N, H, Wd, C = x.shape
        O, Cin, KH, KW = W0.shape
        ncat = 2 * O if gated else O
        pl = _conv_plan(N, H, Wd, C, KH, KW, stride, pad, ncat)
        dx, wbw = None, None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            if pl["dx_implicit"]:
                wbw = _pack_filters(W0, W1, 1, pl["cpad_dx"], pl["Kp_dx"], Cin)
            else:
                wbw = _pack_filters(W0, W1, 0, Cin, pl["Kpc"], ncat)
        sink = ctx.sink
        if sink is not None:
            dW0, db0, dW1, db1 = sink
        else:
            dW0 = torch.empty_like(W0)
            dW1 = torch.empty_like(W1) if gated else None
            db0 = torch.empty((O,), dtype=torch.float32, device=x.device) if has_b0 else None
            db1 = torch.empty((O,), dtype=torch.float32, device=x.device) if has_b1 else None
        ws = _ws(L.exvae_conv2d_bwd_workspace_bytes(N, H, Wd, C, KH, KW, stride, pad, O, 1 if gated else 0), x.device)
        L.check(L.exvae_conv2d_bwd(_p(x), _p(wbw), _p(out), _p(sig), _p(dout), N, H, Wd, C, KH, KW, stride, pad, O,
                                   1 if gated else 0, act, lo, hi, _p(dx), _p(dW0), _p(db0), _p(dW1), _p(db1), _p(ws),
                                   ws.numel(), 1 if sink is not None else 0, _stream()), "conv2d_bwd")
        _count(4 + (0 if dx is None else (1 if pl["dx_implicit"] else 2)))
        if sink is not None:
            _sink_done(sink)
            return dx, None, None, None, None, None, None, None, None, None, None
        return dx, dW0, db0, dW1, db1, None, None, None, None, None, None


def conv2d_gated(x, Wh, bh, Wg, bg, stride: int, pad: int) -> torch.Tensor:
    """GatedConv2d (utils/nn.py:72-95, activation=None): conv_h(x) * sigmoid(conv_g(x)); x, result NHWC."""
    return _Conv2dFn.apply(x, Wh, bh, Wg, bg, int(stride), int(pad), ACT_NONE, 0.0, 0.0, _grad_sink(Wh, bh, Wg, bg))


def conv2d(x, W, b=None, stride: int = 1, pad: int = 0, act: int = ACT_NONE, lo: float = 0.0, hi: float = 0.0):
    """nn.Conv2d + fused activation; x, result NHWC."""
    sk = _grad_sink(W, b)
    return _Conv2dFn.apply(x, W, b, None, None, int(stride), int(pad), int(act), float(lo), float(hi),
                           None if sk is None else (sk[0], sk[1], None, None))


class _WeightNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, g, sink):
        L = lib()
        v, g = _f32(v), _f32(g)
        R = v.shape[0]
        K = v.numel() // R
        w = torch.empty_like(v)
        L.check(L.exvae_weight_norm_fwd(_p(v), _p(g), R, K, _p(w), _stream()), "weight_norm")
        _count(1)
        ctx.save_for_backward(v, g)
        ctx.sink = sink
        return w

    @staticmethod
    def backward(ctx, dw):
        L = lib()
        v, g = ctx.saved_tensors
        dw = _f32(dw)
        R = v.shape[0]
        K = v.numel() // R
        sink = ctx.sink
        dv, dg = sink if sink is not None else (torch.empty_like(v), torch.empty_like(g))
        L.check(L.exvae_weight_norm_bwd(_p(v), _p(g), _p(dw), R, K, _p(dv), _p(dg), 1 if sink is not None else 0, _stream()),
                "weight_norm_bwd")
        _count(1)
        if sink is not None:
            _sink_done(sink)
            return None, None, None
        return dv, dg, None


def weight_norm(v, g) -> torch.Tensor:
    """torch.nn.utils.weight_norm: w = g * v / ||v|| with the norm over everything but dim 0 (models/fully_conv.py:17)."""
    return _WeightNorm.apply(v, g, _grad_sink(v, g))


class _Elu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        L = lib()
        x = _f32(x)
        y = torch.empty_like(x)
        L.check(L.exvae_elu_fwd(_p(x), x.numel(), _p(y), _stream()), "elu")
        _count(1)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        L = lib()
        (y,) = ctx.saved_tensors
        dy = _f32(dy)
        dx = torch.empty_like(y)
        L.check(L.exvae_elu_bwd(_p(y), _p(dy), y.numel(), _p(dx), _stream()), "elu_bwd")
        _count(1)
        return dx


def elu(x) -> torch.Tensor:
    return _Elu.apply(x)


class _Up2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        L = lib()
        x = _f32(x)
        N, H, W, C = x.shape
        y = torch.empty((N, 2 * H, 2 * W, C), dtype=torch.float32, device=x.device)
        L.check(L.exvae_upsample2x_nhwc_fwd(_p(x), N, H, W, C, _p(y), _stream()), "upsample2x")
        _count(1)
        ctx.shape = (N, H, W, C)
        return y

    @staticmethod
    def backward(ctx, dy):
        L = lib()
        N, H, W, C = ctx.shape
        dy = _f32(dy)
        dx = torch.empty((N, H, W, C), dtype=torch.float32, device=dy.device)
        L.check(L.exvae_upsample2x_nhwc_bwd(_p(dy), N, H, W, C, _p(dx), _stream()), "upsample2x_bwd")
        _count(1)
        return dx


def upsample2x(x) -> torch.Tensor:
    """nn.Upsample(scale_factor=2), nearest, NHWC."""
    return _Up2.apply(x)
