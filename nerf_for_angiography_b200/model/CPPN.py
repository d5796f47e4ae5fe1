"""CPPN -- the reference MLP (/root/reference/model/CPPN.py) on hand-written sm_100a kernels.

Same constructor dictionary, same module tree and therefore the same ``state_dict`` keys
(``early_pts_layers.{0,2,..}.{weight,bias}``, ``output_linear.0.*``, ``fourier_coefficients``, ``img1``,
``img2``), same ``forward(x[S,3]) -> [S,1]`` / ``save(filename, info)`` surface.  The configuration the
reference driver enables is implemented (relu, no skip layers, no view directions, pos_enc 'none' or
'fourier'); anything else raises NotImplementedError instead of silently computing something different.

All parameters live in ONE flat fp32 buffer (the nn.Parameters are views into it) so the fused Adam kernel,
the NCCL gradient all-reduce and the weight-packing kernel each touch a single contiguous range.

Extra (optional) key in ``model_definition``: ``'precision'`` in {'bf16', 'fp32'}; default 'bf16' (tcgen05
tensor cores, fp32 accumulate) when the shape is supported by the fused kernel, else 'fp32' (check mode).
"""
import torch
import torch.nn as nn

from .. import ops


class _CPPNForward(torch.autograd.Function):
    """y = MLP(x); backward fills the parameter gradients (no gradient w.r.t. x: the reference's sample
    positions come from the no-grad marcher)."""

    @staticmethod
    def forward(ctx, x, model, sample_kw, *params):
        need_grad = any(ctx.needs_input_grad[3:])   # grad mode is already off inside Function.forward
        prec = model._precision_id
        kp = model._kernel_params()
        packed = model._packed_weights(kp) if prec == ops.PREC_BF16 else None
        kw = sample_kw if sample_kw is not None else {"points": x}
        if need_grad:
            out, saved = ops.mlp_forward(model._desc, kp, packed, ops.OUT_LOGIT, prec, saved=True, **kw)
            ctx.model, ctx.kw, ctx.saved, ctx.packed, ctx.kp = model, kw, saved, packed, kp
        else:
            out = ops.mlp_forward(model._desc, kp, packed, ops.OUT_LOGIT, prec, **kw)
        return out.unsqueeze(-1)

    @staticmethod
    def backward(ctx, grad_out):
        model = ctx.model
        g = grad_out.reshape(-1).contiguous().float()
        flat_grad = model._map_grad_(ops.mlp_backward(model._desc, ctx.kp, ctx.packed, ctx.saved, g, model._precision_id, **ctx.kw))
        grads = [flat_grad[o:o + n].view(shape) if need else None
                 for (o, n, shape), need in zip(model._param_slices, ctx.needs_input_grad[3:])]
        return (None, None, None, *grads)


class CPPN(nn.Module):
    def __init__(self, model_definition: dict) -> None:
        super().__init__()
        self.version = "v0.00"
        self.model_definition = model_definition
        self.device = model_definition['device']
        self.num_early_layers = model_definition['num_early_layers']
        self.num_late_layers = model_definition['num_late_layers']
        self.num_filters = model_definition['num_filters']
        self.num_input_channels = model_definition['num_input_channels']
        self.num_input_channels_views = model_definition['num_input_channels_views']
        self.num_output_channels = model_definition['num_output_channels']
        self.use_bias = model_definition['use_bias']
        self.use_pos_enc = model_definition['pos_enc']
        self.num_img = model_definition['num_img']
        self.mult_img = self.num_img > 1
        self.use_viewdirs = self.num_input_channels_views > 0
        self.store_activations = False
        self.activation_dictionary = {}

        unsupported = []
        if model_definition.get('act_func', 'relu') != 'relu': unsupported.append("act_func != 'relu'")
        if self.num_late_layers != 0: unsupported.append("num_late_layers > 0 (skip connection)")
        if self.use_viewdirs: unsupported.append("view directions")
        if self.num_input_channels != 3 or self.num_output_channels != 1: unsupported.append("channels != (3 -> 1)")
        if not self.use_bias: unsupported.append("use_bias = False")
        if self.use_pos_enc not in ('none', 'fourier', 'barf'): unsupported.append(f"pos_enc = {self.use_pos_enc!r}")
        if self.use_pos_enc == 'fourier' and 'fourier_sigma' not in model_definition: unsupported.append("fourier without fourier_sigma")
        if unsupported:
            raise NotImplementedError("CPPN (B200 hot path) implements the configuration the reference driver uses; unsupported: "
                                      + ", ".join(unsupported))
        self.first_act_func = nn.ReLU()
        self.act_func = nn.ReLU()

        input_features = 3
        self.pos_enc_basis = 0
        if self.use_pos_enc == 'fourier':
            self.pos_enc_basis = int(model_definition['pos_enc_basis'])
            input_features = 3 + 3 * 2 * self.pos_enc_basis
            self.fourier_sigma = model_definition['fourier_sigma']
            self.fourier_coefficients = nn.Parameter(torch.randn([3 * self.pos_enc_basis]) * self.fourier_sigma)
        self._barf = self.use_pos_enc == 'barf'
        if self._barf:
            # BARF (/root/reference/model/CPPN.py:82-94,224-259): features [x | w_k sin(2^k pi x) | w_k cos(2^k pi x)] with a
            # coarse-to-fine mask w_k(alpha).  On the fused kernels this is the Fourier path with FIXED coefficients 2^(k-1)
            # (fl(fl(2 pi x) * 2^(k-1)) == fl(float32(2^k pi) * x) bit for bit: scaling by a power of two commutes with rounding)
            # and the mask folded into the first layer's weight columns; see _kernel_params / _map_grad_.
            self.pos_enc_basis = int(model_definition['pos_enc_basis'])
            input_features = 3 + 3 * 2 * self.pos_enc_basis
            self.k_values = torch.repeat_interleave(torch.arange(0., self.pos_enc_basis), 3)
            self.register_buffer("barf_weights", torch.zeros(3 * self.pos_enc_basis))   # a buffer, not the reference's re-registered Parameter
            self.barf_alpha = 0.0

        H = self.num_filters
        layers = [nn.Linear(input_features, H, bias=True), self.first_act_func]
        for _ in range(self.num_early_layers):
            layers += [nn.Linear(H, H, bias=True), self.act_func]
        self.early_pts_layers = nn.ModuleList(layers)
        self.output_linear = nn.Sequential(nn.Linear(H, 1, bias=True))
        self.img1 = nn.Parameter(torch.tensor([0., 0.], dtype=torch.float))
        self.img2 = nn.Parameter(torch.tensor([0., 0.], dtype=torch.float))

        enc_on = self.use_pos_enc in ('fourier', 'barf') and self.pos_enc_basis > 0
        self._desc = ops.mlp_desc(1 if enc_on else 0, self.pos_enc_basis if enc_on else 0, H, self.num_early_layers)
        self._requested_precision = model_definition.get('precision', None)
        self._flat = None
        self._packed = None
        self._param_slices = []
        self._precision_id = ops.PREC_FP32

    # ------------------------------------------------------------------ flat parameter storage
    def _hot_params(self):
        ps = []
        if self._desc.enc:
            if self._barf:          # fixed frequencies 2^(k-1) occupy the coefficient slots of the flat layout (never trained)
                if getattr(self, "_barf_coef", None) is None or self._barf_coef.device != self.early_pts_layers[0].weight.device:
                    self._barf_coef = (2.0 ** (self.k_values - 1.0)).to(self.early_pts_layers[0].weight.device)
                ps.append(self._barf_coef)
            else:
                ps.append(self.fourier_coefficients)
        for m in self.early_pts_layers:
            if isinstance(m, nn.Linear):
                ps += [m.weight, m.bias]
        ps += [self.output_linear[0].weight, self.output_linear[0].bias]
        return ps

    def _flatten(self):
        """(Re)build the flat fp32 buffer on the parameters' current device and alias every parameter into it."""
        ps = self._hot_params()
        dev = ps[-1].device
        total = sum(p.numel() for p in ps)
        flat = torch.empty(total, dtype=torch.float32, device=dev)
        self._param_slices = []
        off = 0
        for p in ps:
            n = p.numel()
            flat[off:off + n].copy_(p.data.reshape(-1).float())
            if isinstance(p, nn.Parameter):
                p.data = flat[off:off + n].view(p.shape)
            self._param_slices.append((off, n, tuple(p.shape)))
            off += n
        self._flat = flat
        self._packed = None
        if dev.type == 'cuda':
            assert total == ops.mlp_param_count(self._desc), "flat layout mismatch with the C ABI"
            want = self._requested_precision
            supported = ops.mlp_bf16_supported(self._desc)
            if want == 'bf16' and not supported:
                raise NotImplementedError("precision='bf16': this MLP shape is not supported by the tcgen05 kernel")
            self._precision_id = ops.PREC_BF16 if (want == 'bf16' or (want is None and supported)) else ops.PREC_FP32

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._flatten()
        return out

    def _ensure_flat(self):
        ps = self._hot_params()
        if self._flat is None or self._flat.device != ps[-1].device or \
                any(isinstance(p, nn.Parameter) and p.data_ptr() != self._flat.data_ptr() + 4 * o for p, (o, _, _) in zip(ps, self._param_slices)):
            self._flatten()

    def _kernel_params(self):
        """Flat parameters as the kernels see them.  BARF: a copy with the mask w_k folded into the sin/cos columns of the
        first layer (three tiny elementwise ops per call); otherwise the master buffer itself."""
        if not self._barf:
            return self._flat
        nc = 3 * self.pos_enc_basis
        d_in = 3 + 2 * nc
        H = self.num_filters
        eff = self._flat.clone()
        w0 = eff[nc:nc + H * d_in].view(H, d_in)
        w = self.barf_weights.to(eff.device)
        w0[:, 3:3 + nc] *= w
        w0[:, 3 + nc:] *= w
        return eff

    def _map_grad_(self, grad):
        """Gradient w.r.t. the kernel parameters -> gradient w.r.t. the master parameters, in place (BARF: chain rule through
        the folded mask; the fixed frequencies get no gradient)."""
        if self._barf:
            nc = 3 * self.pos_enc_basis
            d_in = 3 + 2 * nc
            H = self.num_filters
            g0 = grad[nc:nc + H * d_in].view(H, d_in)
            w = self.barf_weights.to(grad.device)
            g0[:, 3:3 + nc] *= w
            g0[:, 3 + nc:] *= w
            grad[:nc].zero_()
        return grad

    def update_barf_alpha(self, barf_alpha, type='pts'):
        """/root/reference/model/CPPN.py:236-259 (the mask formula with its quirks: 3.1415 and alpha - k + 1)."""
        if not self._barf or type != 'pts':
            raise NotImplementedError("update_barf_alpha: only the point encoding of a pos_enc='barf' model")
        self.barf_alpha = barf_alpha
        w = []
        for k in self.k_values:
            barf_k = barf_alpha - (k + 1)
            if barf_k < 0:
                w.append(0)
            elif barf_k < 1:
                w.append((1 - torch.cos((barf_alpha - k + 1) * 3.1415)) / 2)
            else:
                w.append(1)
        self.barf_weights.copy_(torch.Tensor(w))

    def _packed_weights(self, kernel_params=None):
        """bf16 UMMA-swizzled weight image, refreshed from the fp32 master parameters on every call."""
        self._packed = ops.mlp_pack(self._desc, self._kernel_params() if kernel_params is None else kernel_params, self._packed)
        return self._packed

    @property
    def precision(self):
        return 'bf16' if self._precision_id == ops.PREC_BF16 else 'fp32'

    def set_precision(self, precision: str):
        if precision not in ('bf16', 'fp32'):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self._requested_precision = precision
        self._ensure_flat()
        self._flatten()

    # ------------------------------------------------------------------ reference surface
    def activations(self, store_activations: bool) -> None:
        if store_activations:
            raise NotImplementedError("activation capture is not available on the fused path")
        self.store_activations = False
        self.activation_dictionary = {}

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("CPPN.forward: input must be a CUDA tensor (there is no CPU fallback)")
        if x.dim() != 2 or x.shape[-1] != 3:
            raise ValueError(f"CPPN.forward expects [S, 3] points, got {tuple(x.shape)}")
        self._ensure_flat()
        x = x.detach().contiguous().float()
        return _CPPNForward.apply(x, self, None, *self._hot_params())

    def forward_samples(self, rays_o, rays_d, ray_idx, t_starts, t_ends) -> torch.Tensor:
        """forward() on the ray-sample midpoints o[idx] + d[idx]*(t0+t1)/2 formed inside the kernel
        (fuses /root/reference/nerf/run_nerf_acc.py:290-294)."""
        self._ensure_flat()
        kw = dict(rays_o=rays_o, rays_d=rays_d, ray_idx=ray_idx, t_starts=t_starts, t_ends=t_ends)
        return _CPPNForward.apply(t_starts, self, kw, *self._hot_params())

    @torch.no_grad()
    def query(self, out_mode, **sample_kw) -> torch.Tensor:
        """no-grad forward with a fused output transform (ops.OUT_LOGIT / OUT_SIGMA / OUT_ALPHA)."""
        self._ensure_flat()
        kp = self._kernel_params()
        packed = self._packed_weights(kp) if self._precision_id == ops.PREC_BF16 else None
        return ops.mlp_forward(self._desc, kp, packed, out_mode, self._precision_id, **sample_kw)

    def save(self, filename: str, training_information: dict) -> None:
        torch.save({'version': self.version, 'parameters': self.model_definition,
                    'training_information': training_information, 'model': self.state_dict()}, f=filename)
