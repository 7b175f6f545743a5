"""Host-side mirror of the reference's model base classes (same method names, signatures and
state_dict keys) over the exvae_b200 CUDA kernels:

    BaseModel   models/BaseModel.py:16-271
    AbsModel    models/AbsModel.py:9-49       (1-level VAE)
    BaseHModel  models/AbsHModel.py:9-106     (2-level VAE)

B200-first differences that do not change results:
  * the training set lives in HBM (``resident``) and exemplars are gathered on the device,
    instead of a CPU gather + 78 MB host->device copy every step (BaseModel.py:247,267);
  * ``log_p_z(sum=True)`` never materialises the [B,C] matrix: distance, mask, normaliser and
    log-sum-exp are one kernel (``ops.prior_lse``), optionally over a range-sharded bank;
  * random draws (eps, exemplar indices) come from a device-side counter-based generator and
    can be injected (``rng_override``) so runs can be replayed against the oracle.
"""
from __future__ import annotations

import math
import os
import weakref
from abc import ABC, abstractmethod
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .distributions import log_bernoulli, log_logistic_256, log_normal_diag, log_normal_standard, pairwise_distance
from .layers import NonLinear, he_init, normal_init


class DeviceRng:
    """Philox4x32-10 stream (seed, call-site subsequence, device counter).  The counter lives in
    device memory ([offset, ticket]) and every drawing kernel advances it itself once all its blocks
    have read it, so a captured CUDA graph draws fresh numbers on every replay."""

    SUB_BERNOULLI, SUB_EXEMPLAR, SUB_EPS = 1, 2, 3

    def __init__(self, seed: int = 0):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self._counter = None

    def counter(self, device) -> torch.Tensor:
        if self._counter is None or self._counter.device != torch.device(device):
            self._counter = torch.zeros(2, dtype=torch.int64, device=device)
        return self._counter

    def bernoulli(self, p):
        c = self.counter(p.device)
        return ops.rng_bernoulli(p, self.seed, c, self.SUB_BERNOULLI, advance=True)

    def normal(self, shape, device, sub=0):
        c = self.counter(device)
        return ops.rng_normal(tuple(shape), self.seed, c, self.SUB_EPS + sub, device, advance=True)

    def randint(self, low, high, n, device):
        c = self.counter(device)
        return ops.rng_randint(low, high, n, self.seed, c, self.SUB_EXEMPLAR, device, advance=True)


def to_nhwc(x, chw):
    """[N, C*H*W] / [N,C,H,W] (the reference's NCHW) -> NHWC for the conv kernels (free when C == 1)."""
    C, H, W = chw
    x = x.reshape(-1, C, H, W)
    return x.reshape(-1, H, W, 1) if C == 1 else x.permute(0, 2, 3, 1).contiguous()


def flat_chw(h):
    """NHWC feature map -> [N, C*H*W] flattened in the reference's (C,H,W) order."""
    N, H, W, C = h.shape
    return h.reshape(N, -1) if C == 1 else h.permute(0, 3, 1, 2).reshape(N, -1)


class PaddedBank(tuple):
    """(centers [cap,D], log_variance, dataset indices [cap]) of the kNN mode with a FIXED capacity cap = B*k: the
    number of selected exemplars is data dependent (torch.unique, models/BaseModel.py:265) and stays on the device
    as ``valid_count`` ([1] int32), so the whole step is graph-capturable; entries beyond ``valid_count`` repeat entry 0 and are
    ignored by the prior kernel."""

    def __new__(cls, items, count):
        self = super().__new__(cls, items)
        self.valid_count = count
        return self

    def valid(self):
        """The reference's variable-length triple (one host sync)."""
        n = int(self.valid_count.item())
        return tuple(t[:n] for t in self)


class BaseModel(nn.Module, ABC):
    knn_graph_capturable = True        # get_approximate_nearest_exemplars has no host sync

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.rng = DeviceRng(getattr(args, "seed", 0))
        self.rng_override: Optional[dict] = None   # {'eps': [tensors...], 'exemplar_indices': tensor}
        self.bank_group = None                      # torch.distributed group when the bank is range-sharded
        self.bank_world, self.bank_rank = 1, 0
        self.grad_sync = None                       # set by distributed.shard_bank: averages the flat gradient buffer
        self.fuse_exemplar_encoder = True           # encode batch + exemplars in ONE pass of the shared trunk
        self.overlap_prior = False                  # run the prior term on a side stream next to the decoder (AbsModel)
        # log-variance head on a side stream next to the mean head: measured gain 4 us of a 790 us step (the library's single
        # side stream serialises the small head's dW behind the large head's), so it stays off by default
        self.parallel_heads = os.environ.get("EXVAE_PARALLEL_HEADS", "0") == "1"
        self._prior_ctx = None                      # (embedding, x_indices) announced by calculate_loss for the side branch
        self._log_p_z_early = None
        self._log_q_early = None                    # log q(z|x) computed together with z in forward()
        self._side_streams = {}
        self._exemplar_prefetch = None              # GraphedTrainStep: persistent fused operand filled one step ahead
        self._zero_rows = {}
        self._resident_cache = {}

        if self.args.prior == 'vampprior':
            self.add_pseudoinputs()
        if self.args.prior == 'exemplar_prior':
            self.prior_log_variance = torch.nn.Parameter(torch.randn((1)))

        if self.args.input_type == 'binary':
            self.p_x_mean = NonLinear(self.args.hidden_size, np.prod(self.args.input_size), activation=nn.Sigmoid())
        elif self.args.input_type in ('gray', 'continuous'):
            self.p_x_mean = NonLinear(self.args.hidden_size, np.prod(self.args.input_size))
            self.p_x_logvar = NonLinear(self.args.hidden_size, np.prod(self.args.input_size),
                                        activation=nn.Hardtanh(min_val=-4.5, max_val=0))
            self.decoder_logstd = torch.nn.Parameter(torch.tensor([0.], requires_grad=True))

        self.create_model(args)
        self.he_initializer()

    def he_initializer(self):
        """models/BaseModel.py:39-44"""
        for m in self.modules():
            if isinstance(m, nn.Linear):
                he_init(m)

    @abstractmethod
    def create_model(self, args):
        pass

    @abstractmethod
    def kl_loss(self, latent_stats, exemplars_embedding, dataset, cache, x_indices):
        pass

    # ------------------------------------------------------------------ data residency
    def resident(self, dataset) -> torch.Tensor:
        """Device-resident copy of ``dataset.tensors[0]`` ([T,P] fp32), uploaded once."""
        src = dataset.tensors[0] if hasattr(dataset, "tensors") else dataset
        dev = self.prior_device()
        if src.is_cuda:
            return src
        # keyed on the tensor OBJECT (weak reference) and its version counter: a new dataset tensor at a recycled
        # address, or an in-place edit of the data (re-binarisation, swapped splits), uploads again
        ent = self._resident_cache.get("entry")
        if ent is not None:
            ref, version, shape, hit = ent
            if ref() is src and version == src._version and shape == tuple(src.shape):
                return hit
        hit = src.to(dev, dtype=torch.float32).contiguous()
        self._resident_cache = {"entry": (weakref.ref(src), src._version, tuple(src.shape), hit)}
        return hit

    def invalidate_resident(self):
        """Drop the device copy of the train set (the next ``resident`` call uploads again)."""
        self._resident_cache = {}

    def prior_device(self):
        return next(self.parameters()).device

    # ------------------------------------------------------------------ loss
    def reconstruction_loss(self, x, x_mean, x_logvar):
        """models/BaseModel.py:54-63"""
        if self.args.input_type == 'binary':
            return log_bernoulli(x, x_mean, dim=1)
        elif self.args.input_type in ('gray', 'continuous'):
            if self.args.use_logit is True:
                return log_normal_diag(x, x_mean, x_logvar, dim=1)
            return log_logistic_256(x, x_mean, x_logvar, dim=1)
        raise Exception('Wrong input type!')

    def calculate_loss(self, x, beta=1., average=False, exemplars_embedding=None, cache=None, dataset=None):
        """models/BaseModel.py:65-77 — returns (loss, RE, KL): scalars if ``average`` else [B]."""
        x, x_indices = x
        zq = None
        self._prior_ctx = self._log_p_z_early = self._log_q_early = None
        if (self.fuse_exemplar_encoder and exemplars_embedding is None and dataset is not None
                and self.args.prior == 'exemplar_prior' and self.args.approximate_prior is False):
            # B200-first: the reference encodes the batch (q_z(x)) and the N exemplars (q_z(.., prior=True))
            # in two passes of the same trunk; here they share one GEMM chain over B+N rows.
            zq, exemplars_embedding = self.q_z_with_exemplars(x, dataset)
        if (self.overlap_prior and exemplars_embedding is not None and self.args.prior == 'exemplar_prior'
                and isinstance(exemplars_embedding, tuple)):
            # the embedding precedes z on this stream: forward() may fork the prior term next to the decoder
            self._prior_ctx = (exemplars_embedding, x_indices)
        x_mean, x_logvar, latent_stats = self.forward(x, zq=zq)
        RE = self.reconstruction_loss(x.reshape(x_mean.shape), x_mean, x_logvar)
        KL = self.kl_loss(latent_stats, exemplars_embedding, dataset, cache, x_indices)
        if average:
            out3 = ops.elbo_reduce(RE, KL, beta, True)
            return out3[0], out3[1], out3[2]
        return ops.elbo_reduce(RE, KL, beta, False), RE, KL

    def _next_eps(self, shape, device, sub=0):
        ro = self.rng_override
        if ro is not None and ro.get('eps'):
            eps = ro['eps'].pop(0)
            assert tuple(eps.shape) == tuple(shape), (eps.shape, shape)
            return eps.to(device)
        return self.rng.normal(shape, device, sub)

    def reparameterize(self, mu, logvar, eps=None, sub=0):
        """models/BaseModel.py:79-82 (eps ~ N(0,1) drawn on the device unless injected)."""
        if eps is None:
            eps = self._next_eps(mu.shape, mu.device, sub)
        return ops.reparameterize(mu, logvar.expand_as(mu), eps)

    def _reparam_with_logq(self, mu, logvar, sub=0):
        """(z, log q(z|x)) in one kernel each way: ``reparameterize`` followed by the ``log_normal_diag(z, mu, logvar)``
        that kl_loss evaluates (models/AbsModel.py:18, models/AbsHModel.py:21,27), bit-identical to the two calls."""
        eps = self._next_eps(mu.shape, mu.device, sub)
        return ops.reparam_logq(mu, logvar.expand_as(mu), eps)

    # ------------------------------------------------------------------ exemplar prior
    def _bank_logvar_row(self, center_log_variance):
        """Row 0 of the bank's log-variance (models/BaseModel.py:101).  When the bank is the learned scalar broadcast
        (q_z(prior=True): a stride-0 view of ``prior_log_variance``) the row is taken from the parameter directly:
        autograd's select + expand backward would otherwise zero-fill a [C,D] tensor and reduce it again."""
        if center_log_variance.dim() != 2:
            return center_log_variance
        plv = getattr(self, "prior_log_variance", None)
        if (plv is not None and center_log_variance.stride() == (0, 0)
                and center_log_variance.data_ptr() == plv.data_ptr()):
            return ops.bcast_scalar(plv, center_log_variance.shape[1])   # [D] row; the gradient sum is one kernel
        return center_log_variance[0, :]

    def _zeros_row(self, P, device):
        """[1,P] zeros (x_logvar of a Bernoulli decoder, models/AbsModel.py:38), allocated once per device."""
        key = (P, str(device))
        z = self._zero_rows.get(key)
        if z is None:
            z = self._zero_rows[key] = torch.zeros(1, P, device=device)
        return z

    def log_p_z_exemplar(self, z, z_indices, exemplars_embedding, test):
        """models/BaseModel.py:98-109 — the [B,C] matrix (materialising, no autograd)."""
        if isinstance(exemplars_embedding, PaddedBank):          # materialising path: the reference's variable-length bank
            exemplars_embedding = exemplars_embedding.valid()
        centers, center_log_variance, center_indices = exemplars_embedding
        lv = center_log_variance[0, :] if center_log_variance.dim() == 2 else center_log_variance
        masked = (test is False) and (self.args.no_mask is False)
        return ops.prior_logprob_matrix(z, centers, lv, z_indices if masked else None,
                                        center_indices if masked else None)

    def _fork_prior(self, z_q):
        """Called by forward() right after z is available: when calculate_loss has announced the exemplar embedding
        (``_prior_ctx``), start log p(z) on a side stream HERE, before the decoder is enqueued.

        The prior term and the decoder are independent chains between z and the loss, and the decoder's GEMMs over B
        rows leave most SMs idle.  Creating the prior ops BEFORE the decoder ops matters for the backward: autograd
        runs nodes in reverse creation order and makes the consumer stream wait for a producer stream as soon as the
        producing node has been issued, so a prior branch created last would be issued first and serialise the whole
        decoder backward behind it (measured: 85 us of an 858 us step).  Only used when the embedding was enqueued
        before z on the main stream (fused batch+exemplar encoder pass, or a bank handed in by the caller).  A capturing
        stream turns the fork into two parallel graph branches; with a range-sharded bank the branch also carries the
        LSE-partial exchange."""
        ctx, self._prior_ctx = self._prior_ctx, None
        self._log_p_z_early = None
        if ctx is None or not z_q.is_cuda:
            return
        exemplars_embedding, x_indices = ctx
        cur = torch.cuda.current_stream()
        side = self._prior_stream()
        centers, clv, cidx = exemplars_embedding
        emb = (centers, self._bank_logvar_row(clv), cidx)     # the parameter's view is taken on the main stream
        c_total = getattr(exemplars_embedding, "c_total", None)
        if c_total is not None:
            from .distributed import ShardedBank
            emb = ShardedBank(emb, c_total)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            ops.take_prior_fwd_event()
            log_p_z = self.log_p_z(z=(z_q, x_indices), exemplars_embedding=emb)
            # with a known upstream gradient the branch goes on into the K1 backward: the loss waits for the forward only
            ev = ops.take_prior_fwd_event()
        self._log_p_z_early = (log_p_z, side, ev)

    def _log_p_z_branch(self, z_q, x_indices, exemplars_embedding):
        """log p(z) of the exemplar prior: the result of the side-stream branch forked in forward(), else computed here."""
        early, self._log_p_z_early = self._log_p_z_early, None
        if early is not None:
            log_p_z, side, ev = early
            cur = torch.cuda.current_stream()
            if ev is not None:
                cur.wait_event(ev)
            else:
                cur.wait_stream(side)
            log_p_z.record_stream(cur)
            return log_p_z
        return self.log_p_z(z=(z_q, x_indices), exemplars_embedding=exemplars_embedding)

    def _known_prior_grad(self, rows):
        """d loss / d log p(z_b) for every row, when the caller of calculate_loss has announced it (``prior_grad_known``,
        set by GraphedTrainStep around its own loss = mean(-RE + beta*KL)): the K1 backward then runs right behind the K1
        forward, next to the decoder.  None (the default) keeps the ordinary autograd backward."""
        g = getattr(self, "prior_grad_known", None)
        if g is None or not self.training or not torch.is_grad_enabled() or g.numel() != rows:
            return None
        return g

    def _aux_stream(self):
        dev = torch.cuda.current_device()
        st = self._side_streams.get((dev, "aux"))
        if st is None:
            st = self._side_streams[(dev, "aux")] = torch.cuda.Stream(device=dev)
        return st

    def _prior_stream(self):
        dev = torch.cuda.current_device()
        st = self._side_streams.get(dev)
        if st is None:
            st = self._side_streams[dev] = torch.cuda.Stream(device=dev)
        return st

    # ------------------------------------------------------------------ VampPrior
    def add_pseudoinputs(self):
        """models/BaseModel.py:130-139 — C learned pseudo-inputs as the weight of a bias-free layer fed the identity."""
        C, P = int(self.args.number_components), int(np.prod(self.args.input_size))
        self.means = NonLinear(C, P, bias=False, activation=nn.Hardtanh(min_val=0.0, max_val=1.0))
        if getattr(self.args, "use_training_data_init", False):
            self.means.linear.weight.data = self.args.pseudoinputs_mean
        else:
            normal_init(self.means.linear, getattr(self.args, "pseudoinputs_mean", -0.05),
                        getattr(self.args, "pseudoinputs_std", 0.01))
        self.register_buffer("idle_input", torch.eye(C, C), persistent=False)

    def pseudo_embedding(self):
        """(z_p_mean [C,D], z_p_logvar [C,D]) = q_z(means(idle_input), prior=True)  (BaseModel.py:86-88)"""
        return self.q_z(self.means(self.idle_input), prior=True)

    def log_p_z_vampprior(self, z, exemplars_embedding, sum=True):
        """models/BaseModel.py:84-96 (+ the log-sum-exp of :123-125 when ``sum``)."""
        if exemplars_embedding is None:
            z_p_mean, z_p_logvar = self.pseudo_embedding()
        else:
            z_p_mean, z_p_logvar = exemplars_embedding[0], exemplars_embedding[1]
        if not sum:
            return ops.vamp_logprob_matrix(z, z_p_mean, z_p_logvar)
        return ops.vamp_lse(z, z_p_mean, z_p_logvar)

    def log_p_z(self, z, exemplars_embedding, sum=True, test=None):
        """models/BaseModel.py:111-128"""
        z, z_indices = z
        if test is None:
            test = not self.training
        if self.args.prior == 'standard':
            return log_normal_standard(z, dim=1)
        elif self.args.prior == 'vampprior':
            return self.log_p_z_vampprior(z, exemplars_embedding, sum=sum)
        elif self.args.prior == 'exemplar_prior':
            if not sum:
                return self.log_p_z_exemplar(z, z_indices, exemplars_embedding, test)
            centers, center_log_variance, center_indices = exemplars_embedding
            lv = self._bank_logvar_row(center_log_variance)
            masked = (test is False) and (self.args.no_mask is False) and z_indices is not None
            c_total = getattr(exemplars_embedding, "c_total", None)
            if c_total is not None and self.bank_group is not None:     # range-sharded bank (distributed.py)
                return ops.prior_lse_sharded(z, centers, lv, z_indices if masked else None,
                                             center_indices if masked else None, c_total, self.bank_group,
                                             g_known=self._known_prior_grad(z.shape[0] * self.bank_world))
            return ops.prior_lse(z, centers, lv, z_indices if masked else None, center_indices if masked else None,
                                 c_valid=getattr(exemplars_embedding, "valid_count", None),
                                 g_known=self._known_prior_grad(z.shape[0]))
        raise Exception('Wrong name of the prior!')

    # ------------------------------------------------------------------ generation helpers
    def generate_z(self, N=25, dataset=None):
        """models/BaseModel.py:152-166"""
        dev = self.prior_device()
        if self.args.prior == 'standard':
            return self.rng.normal((N, self.args.z1_size), dev)
        if self.args.prior == 'vampprior':
            means = self.means(self.idle_input)[0:N]
            mean, logvar = self.q_z(means)
            return self.reparameterize(mean, logvar)
        rand_indices = self.rng.randint(0, self.args.training_set_size, N, dev)
        exemplars = ops.gather_rows(self.resident(dataset), rand_indices)
        mean, logvar = self.q_z(exemplars, prior=True)
        return self.reparameterize(mean, logvar)

    def reference_based_generation_z(self, N=25, reference_image=None):
        """models/BaseModel.py:168-174"""
        pseudo, log_var = self.q_z(reference_image.to(self.prior_device()), prior=True)
        pseudo = pseudo.unsqueeze(1).expand(-1, N, -1).reshape(-1, pseudo.shape[-1])
        log_var = log_var[0].unsqueeze(0).expand(len(pseudo), -1)
        z = self.reparameterize(pseudo, log_var)
        return z.reshape(-1, N, pseudo.shape[1])

    def reconstruct_x(self, x):
        x_reconstructed, _, _ = self.forward(x)
        return x_reconstructed

    def generate_x(self, N=25, dataset=None):
        return self.generate_x_from_z(self.generate_z(N=N, dataset=dataset))

    def reference_based_generation_x(self, N=25, reference_image=None):
        z = self.reference_based_generation_z(N=N, reference_image=reference_image)
        return self.generate_x_from_z(z.reshape(-1, z.shape[-1]))

    def logit_inverse(self, x):
        lambd = self.args.lambd
        return (torch.sigmoid(x) - lambd) / (1 - 2 * lambd)

    # ------------------------------------------------------------------ encoder over a bank
    def _is_conv(self):
        return 'conv' in self.args.model_name

    def _trunk(self, x):
        """q_z_layers over [R, P] rows; conv stacks run NHWC and return the feature map the heads expect."""
        if self._is_conv():
            h = self.q_z_layers(to_nhwc(x, self.args.input_size))
            return flat_chw(h) if self.args.model_name == 'convhvae_2level' else h
        return self.q_z_layers(x)

    def _head(self, module, h):
        out = module(h)
        return flat_chw(out) if out.dim() == 4 else out

    def q_z(self, x, prior=False):
        """models/BaseModel.py:205-221 — returns (mean [R,D], logvar [R,D]).  With ``prior=True``
        under the exemplar prior the log-variance is the learned scalar broadcast (a stride-0
        view, no [R,D] tensor is written)."""
        R = x.shape[0]
        h = self._trunk(x)
        z_q_mean = self._head(self.q_z_mean, h)
        if prior is True and self.args.prior == 'exemplar_prior':
            z_q_logvar = self.prior_log_variance.expand(R, self.args.z1_size)
        else:
            z_q_logvar = self._head(self.q_z_logvar, h)
        return z_q_mean.reshape(-1, self.args.z1_size), z_q_logvar.reshape(-1, self.args.z1_size)

    def q_z_with_exemplars(self, x, dataset):
        """(q_z(x), get_exemplar_set(...)) of models/BaseModel.py:205-221,243-248 computed together:
        exemplar rows are gathered straight behind the batch rows and the encoder trunk runs once."""
        dev = x.device
        B = x.shape[0]
        P = int(np.prod(self.args.input_size))
        pf = self._exemplar_prefetch
        if pf is not None and pf["B"] == B and pf["rows"].shape[1] == P and pf["rows"].device == dev:
            # exemplar rows (and their indices) of THIS step were drawn and gathered behind the previous step's backward
            # (prefetch_exemplars): only the batch rows are copied in
            rows, exemplars_indices = pf["rows"], pf["idx"]
            n = exemplars_indices.numel()
            rows[:B].copy_(x.reshape(B, P))
        else:
            exemplars_indices = self._exemplar_indices(dev)
            n = exemplars_indices.numel()
            rows = torch.empty((B + n, P), dtype=torch.float32, device=dev)
            rows[:B].copy_(x.reshape(B, P))
            ops.gather_rows(self.resident(dataset), exemplars_indices, out=rows[B:])
        h = self._trunk(rows)
        h, h_batch = ops.shared_rows(h, B)          # mean head: all rows; log-variance head: batch rows only
        # The two heads only share their input: the small one (B rows, a few CTAs) runs on a side stream next to the
        # large one, forward AND backward (autograd replays a node on the stream of its forward), instead of in front
        # of / behind it on the step's critical path.  A capturing stream turns this into two parallel graph branches.
        aux = self._aux_stream() if (h.is_cuda and self.parallel_heads) else None
        if aux is not None:
            cur = torch.cuda.current_stream()
            aux.wait_stream(cur)
            with torch.cuda.stream(aux):
                z_q_logvar = self._head(self.q_z_logvar, h_batch).reshape(B, -1)
            h_batch.record_stream(aux)
            mean_all = self._head(self.q_z_mean, h).reshape(B + n, -1)
            cur.wait_stream(aux)
            z_q_logvar.record_stream(cur)
        else:
            mean_all = self._head(self.q_z_mean, h).reshape(B + n, -1)
            z_q_logvar = self._head(self.q_z_logvar, h_batch).reshape(B, -1)
        ex_logvar = self.prior_log_variance.expand(n, self.args.z1_size)
        mean_batch, mean_bank = ops.split_rows(mean_all, B)
        exemplar_set = (mean_bank, ex_logvar, exemplars_indices)
        if self.bank_group is not None:
            from .distributed import ShardedBank
            exemplar_set = ShardedBank(exemplar_set, self.args.number_components)
        return (mean_batch, z_q_logvar), exemplar_set

    def exemplar_count(self):
        """Exemplars this rank draws per step (models/BaseModel.py:245; its share of a range-sharded bank)."""
        from .distributed import shard_range
        lo, hi = shard_range(self.args.number_components, self.bank_world, self.bank_rank)
        return hi - lo

    @torch.no_grad()
    def prefetch_exemplars(self, pf, dataset):
        """Draw the NEXT step's exemplar indices (BaseModel.py:245) and gather their rows (:247) into the persistent
        fused operand ``pf = {"rows": [B+n, P], "idx": [n], "B": B}``.  GraphedTrainStep runs this behind the backward of
        the current step, next to the optimizer: the 157 MB gather is HBM bound and nothing else of the step is."""
        dev = pf["rows"].device
        pf["idx"].copy_(self._exemplar_indices(dev))
        ops.gather_rows(self.resident(dataset), pf["idx"], out=pf["rows"][pf["B"]:])

    def cache_z(self, dataset, prior=True, cuda=True):
        """models/BaseModel.py:223-241 — embed the whole (resident) dataset in chunks of 10 000."""
        data = self.resident(dataset)
        cached_z, cached_log_var = [], []
        step = 10000
        for i in range(math.ceil(data.shape[0] / step)):
            m, lv = self.q_z(data[i * step:(i + 1) * step], prior=prior)
            cached_z.append(m)
            cached_log_var.append(lv)
        return torch.cat(cached_z, dim=0), torch.cat(cached_log_var, dim=0)

    def _exemplar_indices(self, device):
        """N dataset indices drawn with replacement (BaseModel.py:245); with a range-sharded bank
        this rank draws (or is handed) only its own N/G of them."""
        from .distributed import shard_range
        n = self.args.number_components
        lo, hi = shard_range(n, self.bank_world, self.bank_rank)
        ro = self.rng_override
        if ro is not None and ro.get('exemplar_indices') is not None:
            idx = ro['exemplar_indices'].to(device).reshape(-1)
            return idx[lo:hi] if (self.bank_world > 1 and idx.numel() == n) else idx
        return self.rng.randint(0, self.args.training_set_size, hi - lo, device)

    def get_exemplar_set(self, z_mean, z_log_var, dataset, cache, x_indices):
        """models/BaseModel.py:243-254"""
        if self.args.approximate_prior is False:
            dev = z_mean.device
            exemplars_indices = self._exemplar_indices(dev)
            exemplars = ops.gather_rows(self.resident(dataset), exemplars_indices)
            exemplars_z, log_variance = self.q_z(exemplars, prior=True)
            if self.bank_group is not None:
                from .distributed import ShardedBank
                return ShardedBank((exemplars_z, log_variance, exemplars_indices), self.args.number_components)
            return (exemplars_z, log_variance, exemplars_indices)
        return self.get_approximate_nearest_exemplars(z=(z_mean, z_log_var, x_indices), dataset=dataset, cache=cache)

    def get_approximate_nearest_exemplars(self, z, cache, dataset):
        """models/BaseModel.py:256-271 — kNN exemplar selection against the cached bank.
        Cache rows are refreshed in place (no autograd through the cache: the reference detaches
        it after every step, utils/training.py:45-46, and only the re-encoded exemplars carry
        gradient)."""
        z, _, indices = z
        dev = z.device
        exemplars_indices = self._exemplar_indices(dev)
        cached_z, cached_log_variance = cache
        ops.scatter_rows_(cached_z, indices.reshape(-1), z.detach())
        sub_cache = ops.gather_rows(cached_z, exemplars_indices)
        nearest_indices, _ = ops.knn_topk(z.detach(), sub_cache, int(self.args.approximate_k))
        # torch.unique (BaseModel.py:265) has a data-dependent size: keep the fixed upper bound B*k and a device-side
        # count instead (no host sync, so the step can be captured in a CUDA graph); the tail repeats entry 0
        nearest, count = ops.unique_positions(nearest_indices, sub_cache.shape[0])
        exemplars_indices = ops.gather_index(exemplars_indices, nearest)
        exemplars = ops.gather_rows(self.resident(dataset), exemplars_indices)
        exemplars_z, log_variance = self.q_z(exemplars, prior=True)
        ops.scatter_rows_(cached_z, exemplars_indices, exemplars_z.detach())
        return PaddedBank((exemplars_z, log_variance, exemplars_indices), count)


class AbsModel(BaseModel):
    """models/AbsModel.py:9-49"""

    def kl_loss(self, latent_stats, exemplars_embedding, dataset, cache, x_indices):
        z_q, z_q_mean, z_q_logvar = latent_stats
        if exemplars_embedding is None and self.args.prior == 'exemplar_prior':
            exemplars_embedding = self.get_exemplar_set(z_q_mean, z_q_logvar, dataset, cache, x_indices)
        log_p_z = self._log_p_z_branch(z_q, x_indices, exemplars_embedding)
        log_q_z, self._log_q_early = self._log_q_early, None     # computed together with z in forward()
        if log_q_z is None:
            log_q_z = log_normal_diag(z_q, z_q_mean, z_q_logvar, dim=1)
        return ops.lincomb((-1.0, 1.0), log_p_z, log_q_z)     # -(log_p_z - log_q_z)

    def generate_x_from_z(self, z, with_reparameterize=True):
        generated_x, _ = self.p_x(z)
        if getattr(self.args, "use_logit", False) is True:
            return self.logit_inverse(generated_x)
        return generated_x

    def p_x(self, z):
        """models/AbsModel.py:31-42"""
        P = int(np.prod(self.args.input_size))
        if self._is_conv():
            from ._lib import ACT_HARDTANH
            from .layers import WNConv2d
            hw = self.args.input_size[1] // 4
            z = to_nhwc(z, (self.bottleneck, hw, hw))
            h = self.p_x_layers(z)
            if self.args.input_type != 'binary' and self.args.use_logit is False and isinstance(self.p_x_mean, WNConv2d):
                x_mean = self.p_x_mean(h, ACT_HARDTANH, 0. + 1. / 512., 1. - 1. / 512.)   # clamp fused in the epilogue
            else:
                x_mean = self.p_x_mean(h)
            x_mean = flat_chw(x_mean)
        else:
            h = self.p_x_layers(z)
            if self.args.input_type != 'binary' and self.args.use_logit is False:
                from ._lib import ACT_HARDTANH       # the clamp of models/AbsModel.py:36 fused in the GEMM epilogue
                lin = self.p_x_mean.linear
                x_mean = ops.linear(h, lin.weight, lin.bias, ACT_HARDTANH, 0. + 1. / 512., 1. - 1. / 512.)
            else:
                x_mean = self.p_x_mean(h)
        if self.args.input_type == 'binary':
            x_logvar = self._zeros_row(P, x_mean.device)
        else:
            x_logvar = self.decoder_logstd.expand(x_mean.shape[0], P)
        return x_mean.reshape(-1, P), x_logvar.reshape(-1, P)

    def forward(self, x, label=0, num_categories=10, zq=None):
        z_q_mean, z_q_logvar = self.q_z(x) if zq is None else zq
        z_q, self._log_q_early = self._reparam_with_logq(z_q_mean, z_q_logvar)
        z_dec, z_q = ops.fanout(z_q, 2)           # consumers: decoder | prior term
        self._fork_prior(z_q)
        x_mean, x_logvar = self.p_x(z_dec)
        return x_mean, x_logvar, (z_q, z_q_mean, z_q_logvar)


class BaseHModel(BaseModel):
    """models/AbsHModel.py:9-106"""

    def kl_loss(self, latent_stats, exemplars_embedding, dataset, cache, x_indices):
        z1_q, z1_q_mean, z1_q_logvar, z2_q, z2_q_mean, z2_q_logvar, z1_p_mean, z1_p_logvar = latent_stats
        if exemplars_embedding is None and self.args.prior == 'exemplar_prior':
            exemplars_embedding = self.get_exemplar_set(z2_q_mean, z2_q_logvar, dataset, cache, x_indices)
        D1, D2 = self.args.z1_size, self.args.z2_size
        early, self._log_q_early = self._log_q_early, None       # (log_q_z2, log_q_z1) computed with z2 / z1 in forward()
        log_p_z1 = log_normal_diag(z1_q.view(-1, D1), z1_p_mean.view(-1, D1), z1_p_logvar.view(-1, D1), dim=1)
        log_q_z1 = early[1] if early is not None else \
            log_normal_diag(z1_q.view(-1, D1), z1_q_mean.view(-1, D1), z1_q_logvar.view(-1, D1), dim=1)
        log_p_z2 = self._log_p_z_branch(z2_q, x_indices, exemplars_embedding)
        log_q_z2 = early[0] if early is not None else \
            log_normal_diag(z2_q.view(-1, D2), z2_q_mean.view(-1, D2), z2_q_logvar.view(-1, D2), dim=1)
        return ops.lincomb((-1.0, -1.0, 1.0, 1.0), log_p_z1, log_p_z2, log_q_z1, log_q_z2)

    def generate_x_from_z(self, z, with_reparameterize=True):
        z1_mean, z1_logvar = self.p_z1(z)
        z1 = self.reparameterize(z1_mean, z1_logvar) if with_reparameterize else z1_mean
        generated_xs, _ = self.p_x(z1.view(-1, self.args.z1_size), z.view(-1, self.args.z2_size))
        return generated_xs

    def p_z1(self, z2):
        z2 = self.p_z1_layers_z2(z2)
        return self.p_z1_mean(z2), self.p_z1_logvar(z2)

    def q_z1(self, x, z2):
        if self._is_conv():
            x = flat_chw(self.q_z1_layers_x(to_nhwc(x, self.args.input_size)))
        else:
            x = self.q_z1_layers_x(x)
        z2 = self.q_z1_layers_z2(z2)
        h = ops.concat_cols(x, z2)
        h = self.q_z1_layers_joint(h)
        return self.q_z1_mean(h), self.q_z1_logvar(h)

    def p_x(self, z1, z2, x=None):
        z1 = self.p_x_layers_z1(z1)
        z2 = self.p_x_layers_z2(z2)
        h = ops.concat_cols(z1, z2)
        conv = 'convhvae_2level' in self.args.model_name
        if conv:
            h = self.p_x_layers_joint_pre(h)
            h = to_nhwc(h, self.args.input_size)
        h_decoder = self.p_x_layers_joint(h)
        x_mean = self.p_x_mean(h_decoder)
        P = int(np.prod(self.args.input_size))
        if conv:
            x_mean = flat_chw(x_mean)
        if self.args.input_type == 'binary':
            x_logvar = self._zeros_row(P, x_mean.device)
        else:
            x_mean = torch.clamp(x_mean, min=0. + 1. / 512., max=1. - 1. / 512.)
            x_logvar = self.p_x_logvar(h_decoder)
            if conv:
                x_logvar = flat_chw(x_logvar)
        return x_mean, x_logvar

    def forward(self, x, zq=None):
        z2_q_mean, z2_q_logvar = self.q_z(x) if zq is None else zq
        z2_q, log_q_z2 = self._reparam_with_logq(z2_q_mean, z2_q_logvar, sub=0)
        z2_a, z2_b, z2_c, z2_q = ops.fanout(z2_q, 4)     # consumers: q(z1|x,z2) | p(z1|z2) | p(x|z1,z2) | prior term
        self._fork_prior(z2_q)
        z1_q_mean, z1_q_logvar = self.q_z1(x, z2_a)
        z1_q, log_q_z1 = self._reparam_with_logq(z1_q_mean, z1_q_logvar, sub=1)
        z1_a, z1_q = ops.fanout(z1_q, 2)                 # consumers: p(x|z1,z2) | log p(z1|z2)
        self._log_q_early = (log_q_z2, log_q_z1)
        z1_p_mean, z1_p_logvar = self.p_z1(z2_b)
        x_mean, x_logvar = self.p_x(z1_a, z2_c)
        return x_mean, x_logvar, (z1_q, z1_q_mean, z1_q_logvar, z2_q, z2_q_mean, z2_q_logvar, z1_p_mean, z1_p_logvar)
