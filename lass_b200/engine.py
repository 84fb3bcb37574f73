"""Host-side driver of the whole-model C-ABI entry (``lass_resunet30_*`` in ``include/lass_b200.h``).

* packs the module's parameters into the kernel layouts (BatchNorm folded to scale/shift, FiLM linears concatenated
  into one GEMM with the BN shift folded into its bias, conv weights to (taps, cout, cin) bf16, shortcuts to fp16,
  DFT basis split hi/lo) — re-packed automatically when any parameter's version counter changes;
* caches one plan (tensor maps + launch list + workspace) per (batch, length);
* runs the forward on the current CUDA stream.  No CPU path: inputs must be CUDA tensors.
"""
import ctypes
import weakref

import torch

from . import _cabi, packing

_ENC = ("encoder_block1", "encoder_block2", "encoder_block3", "encoder_block4", "encoder_block5", "encoder_block6",
        "conv_block7a")
_DEC = ("decoder_block1", "decoder_block2", "decoder_block3", "decoder_block4", "decoder_block5", "decoder_block6")
BN_EPS = 1e-5
# Batches up to this many samples (batch x length) are replayed from a CUDA graph captured per (batch, length): at batch 1 the
# 36 launches + 4 tensor-map encodes of a forward cost more host time than the kernels take on the GPU.
GRAPH_MAX_SAMPLES = 4 * 160000


def _dev_key(device):
    device = torch.device(device)
    return (device.type, device.index if device.index is not None else torch.cuda.current_device())


def fold_bn(bn: torch.nn.BatchNorm2d):
    """eval-mode BatchNorm -> (scale, shift): y = scale * x + shift."""
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    shift = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    return scale, shift


def film_sites(base):
    """[(bn module, FiLM layer name, channel count)] in the row order of ``include/lass_b200.h``."""
    sites = []
    for name in _ENC:
        cb = getattr(base, name).conv_block1
        sites.append((cb.bn1, "%s->conv_block1->beta1" % name))
        sites.append((cb.bn2, "%s->conv_block1->beta2" % name))
    for name in _DEC:
        blk = getattr(base, name)
        sites.append((blk.bn1, "%s->beta1" % name))
        sites.append((blk.conv_block2.bn1, "%s->conv_block2->beta1" % name))
        sites.append((blk.conv_block2.bn2, "%s->conv_block2->beta2" % name))
    return sites


class _Plan:
    def __init__(self, handle, workspace, B, L):
        self.handle, self.workspace, self.B, self.L = handle, workspace, B, L
        self.graph = None             # (torch.cuda.CUDAGraph, static mixture, static condition / shift, static output)

    def __del__(self):
        try:
            if self.handle:
                _cabi.load().lass_resunet30_plan_destroy(self.handle)
        except Exception:
            pass


class Engine:
    """One engine per (base module, FiLM module or None).  The engine does not keep its modules alive (weak references:
    ``models.resunet._ENGINES`` is keyed weakly by the base module, so dropping the model frees the packed weights, plans
    and workspace).  Single-stream use: all plans of an engine share ONE workspace, so forwards of the same engine must be
    issued on one stream at a time (what the reference's one-thread-per-process callers do)."""

    def __init__(self, base, film):
        self._base_ref = weakref.ref(base)
        self._film_ref = weakref.ref(film) if film is not None else None
        self._packed = None
        self._packed_key = None
        self._plans = {}
        self._tensors = None
        self.stft_precision_mode = 0
        self.use_graphs = True

    @property
    def base(self):
        b = self._base_ref()
        if b is None:
            raise RuntimeError("the module this engine was built for no longer exists")
        return b

    @property
    def film(self):
        return self._film_ref() if self._film_ref is not None else None

    # ------------------------------------------------------------------ packing
    def invalidate(self):
        """Forget the cached parameter list (call after REPLACING parameter objects by hand; load_state_dict, .to() and
        in-place updates are detected without it)."""
        self._tensors = None

    def _version_key(self, device):
        # (data_ptr, version) of every parameter / buffer: in-place updates bump the version, .to() / load_state_dict(assign)
        # change the pointer.  Walking the module tree costs ~0.8 ms per call, so the tensor list itself is cached and
        # rebuilt when the module's structure epoch changes (models/resunet.py bumps it in _apply / load_state_dict hooks).
        base, film = self.base, self.film
        epoch = (getattr(base, "_lass_epoch", 0), getattr(film, "_lass_epoch", 0) if film is not None else 0)
        if self._tensors is None or self._tensors[0] != epoch:
            mods = [base] + ([film] if film is not None else [])
            self._tensors = (epoch, [t for m in mods for t in list(m.parameters()) + list(m.buffers())])
        return (_dev_key(device), tuple([(t.data_ptr(), t._version) for t in self._tensors[1]]))

    def _pack(self, device):
        base = self.base
        if base.input_channels != 1 or base.output_channels != 1:
            raise NotImplementedError("the B200 path implements input_channels == output_channels == 1 "
                                      "(reference config/audiosep_base.yaml:25-28)")
        lib = _cabi.load()
        keep = {}

        def dev(t, dtype=torch.float32):
            return t.detach().to(device=device, dtype=dtype).contiguous()

        w = _cabi.ResUNet30Weights()
        w.n_fft, w.hop = base.window_size, base.hop_size
        hi, lo = packing.pack_stft_basis(dev(base.stft.conv_real.weight), dev(base.stft.conv_imag.weight))
        keep["basis"] = (hi, lo)
        win, tw = packing.istft_tables(base.window_size, device=device)
        keep["istft"] = (win, tw)
        # K5 computes the inverse DFT with an FFT and this analytic Hann window instead of multiplying by the module's frozen
        # istft.conv_real / conv_imag matrices; a checkpoint whose matrices are NOT the analytic ones (another window, a
        # trained ISTFT) must not be separated silently with the wrong synthesis -- compare a few rows and fail loudly
        n = base.window_size
        rows = torch.tensor([0, 1, n // 4 + 3, n // 2, n - 1], device=device)
        ang = 2.0 * torch.pi * (rows.double()[:, None] * torch.arange(n, device=device, dtype=torch.float64)[None, :] % n) / n
        want_re = (torch.cos(ang) * win.double()[None, :] / n).float()        # (bins x in rows, samples y)
        got = dev(base.istft.conv_real.weight)[:, :, 0][:, rows].t()          # weight[y, x, 0] = cos(2 pi x y / n) / n * w[y]
        if float((got - want_re).abs().max()) > 64e-6 / n or \
                float((dev(base.istft.ola_window) - win * win).abs().max()) > 1e-6:
            raise ValueError("istft.conv_real / ola_window differ from the analytic periodic-Hann inverse DFT that kernel K5 "
                             "implements (reference torchlibrosa ISTFT); this checkpoint cannot run on the fused iSTFT")
        s0, b0 = fold_bn(base.bn0)
        keep["bn0"] = (dev(s0), dev(b0))
        keep["pre"] = (dev(base.pre_conv.weight.reshape(-1)), dev(base.pre_conv.bias))
        # FiLM GEMM + folded BN
        sites = film_sites(base)
        rows = lib.lass_resunet30_film_rows()
        scales, biases, weights = [], [], []
        cond = None
        for i, (bn, film_name) in enumerate(sites):
            assert lib.lass_resunet30_film_offset(i) == sum(x.numel() for x in scales), "FiLM row order mismatch"
            sc, sh = fold_bn(bn)
            scales.append(dev(sc))
            if self.film is not None:
                lin = getattr(self.film, film_name)
                cond = lin.in_features
                weights.append(dev(lin.weight))
                biases.append(dev(lin.bias) + dev(sh))
            else:
                biases.append(dev(sh))
        act_scale = torch.cat(scales)
        film_b = torch.cat(biases)
        assert act_scale.numel() == rows
        if self.film is not None:
            film_w = torch.cat(weights, dim=0).contiguous()
        else:
            cond = 1
            film_w = torch.zeros(rows, 1, device=device)
        keep["film"] = (film_w, film_b, act_scale)
        w.condition_size, w.film_rows = cond, rows
        w.stft_basis_hi, w.stft_basis_lo = hi.data_ptr(), lo.data_ptr()
        w.istft_window, w.istft_twiddle = win.data_ptr(), tw.data_ptr()
        w.bn0_scale, w.bn0_shift = keep["bn0"][0].data_ptr(), keep["bn0"][1].data_ptr()
        w.pre_w, w.pre_b = keep["pre"][0].data_ptr(), keep["pre"][1].data_ptr()
        w.film_w, w.film_b, w.act_scale = film_w.data_ptr(), film_b.data_ptr(), act_scale.data_ptr()

        def block_weights(cb, cin, cout):
            c1 = packing.pack_conv_weight(dev(cb.conv1.weight), torch.bfloat16)
            c2 = packing.pack_conv_weight(dev(cb.conv2.weight), torch.bfloat16)
            if cb.is_shortcut:
                sc = packing.pack_conv_weight(dev(cb.shortcut.weight), torch.float16)
                sb = dev(cb.shortcut.bias)
            else:
                sc = torch.eye(cout, cin, device=device, dtype=torch.float16).reshape(1, cout, cin).contiguous()
                sb = None
            return c1, c2, sc, sb

        for k, name in enumerate(_ENC):
            cb = getattr(base, name).conv_block1
            c1, c2, sc, sb = block_weights(cb, cb.conv1.in_channels, cb.conv1.out_channels)
            keep["enc%d" % k] = (c1, c2, sc, sb)
            w.enc[k].conv1_w, w.enc[k].conv2_w, w.enc[k].sc_w = c1.data_ptr(), c2.data_ptr(), sc.data_ptr()
            w.enc[k].sc_b = sb.data_ptr() if sb is not None else None
        for j, name in enumerate(_DEC):
            blk = getattr(base, name)
            up = packing.pack_convT_weight(dev(blk.conv1.weight), torch.bfloat16)
            cb = blk.conv_block2
            c1, c2, sc, sb = block_weights(cb, cb.conv1.in_channels, cb.conv1.out_channels)
            keep["dec%d" % j] = (up, c1, c2, sc, sb)
            w.dec[j].up_w, w.dec[j].conv1_w, w.dec[j].conv2_w = up.data_ptr(), c1.data_ptr(), c2.data_ptr()
            w.dec[j].sc_w = sc.data_ptr()
            w.dec[j].sc_b = sb.data_ptr() if sb is not None else None
        keep["after"] = (dev(base.after_conv.weight.reshape(3, 32)), dev(base.after_conv.bias))
        w.after_w, w.after_b = keep["after"][0].data_ptr(), keep["after"][1].data_ptr()
        w.dxn_mask = 0
        keep["struct"] = w
        keep["bn_shift_rows"] = None
        return keep

    def _get_packed(self, device):
        key = self._version_key(device)
        if self._packed is None or key != self._packed_key:
            self._packed = self._pack(device)
            self._packed_key = key
            self._plans = {}          # plans hold pointers into the packed weights
        return self._packed

    def release(self):
        """Free the packed weights, plans (and their CUDA graphs) and the workspace; the next forward rebuilds them."""
        self._plans = {}
        self._packed = None
        self._packed_key = None
        self._ws = None

    # ------------------------------------------------------------------ plans
    def _get_plan(self, B, L, device):
        packed = self._get_packed(device)
        plan = self._plans.get((B, L))
        if plan is not None:
            return plan
        lib = _cabi.load()
        need = lib.lass_resunet30_workspace_bytes(B, L, self.base.window_size, self.base.hop_size)
        if need == 0:
            raise ValueError("unsupported geometry: B=%d L=%d n_fft=%d hop=%d (need L > n_fft/2, hop %% 8 == 0)"
                             % (B, L, self.base.window_size, self.base.hop_size))
        # one workspace shared by all plans of this engine (plans run one at a time on a stream)
        ws = getattr(self, "_ws", None)
        if ws is None or ws.numel() < need or _dev_key(ws.device) != _dev_key(device):
            self._plans = {}
            ws = torch.empty(need + 1024, dtype=torch.uint8, device=device)
            self._ws = ws
        base_ptr = (ws.data_ptr() + 1023) // 1024 * 1024
        handle = ctypes.c_void_p()
        _cabi.check(lib.lass_resunet30_plan_create(ctypes.byref(packed["struct"]), B, L, base_ptr,
                                                   ws.numel() - (base_ptr - ws.data_ptr()), ctypes.byref(handle)))
        plan = _Plan(handle, ws, B, L)
        if len(self._plans) > 8:
            self._plans.pop(next(iter(self._plans)))
        self._plans[(B, L)] = plan
        return plan

    # ------------------------------------------------------------------ forward
    def _check_inputs(self, mixtures):
        if self.base.training:
            raise RuntimeError(
                "the inference engine runs eval-mode BatchNorm; in .train() call the ResUNet30 module itself "
                "(batch-statistics forward + backward: lass_b200/training.py) or switch to .eval() first")
        if not mixtures.is_cuda:
            raise RuntimeError("lass_b200 has no CPU path: move the module and its inputs to a CUDA device")
        if mixtures.dim() != 3 or mixtures.shape[1] != 1:
            raise ValueError("mixture must be (batch, 1, samples); got %s" % (tuple(mixtures.shape),))

    def _launch(self, plan, x_ptr, cond_ptr, shift_ptr, out_ptr):
        _cabi.check(_cabi.load().lass_resunet30_forward(plan.handle, x_ptr, cond_ptr, shift_ptr, out_ptr,
                                                        self.stft_precision_mode, torch.cuda.current_stream().cuda_stream))

    def _run(self, plan, x, cond, shift, device):
        """One forward over contiguous fp32 CUDA tensors (cond XOR shift); returns a fresh (B, 1, L) tensor."""
        B, L = plan.B, plan.L
        aux = cond if cond is not None else shift
        graphable = (self.use_graphs and B * L <= GRAPH_MAX_SAMPLES and not torch.cuda.is_current_stream_capturing())
        if not graphable:
            out = torch.empty(B, 1, L, dtype=torch.float32, device=device)
            self._launch(plan, x.data_ptr(), cond.data_ptr() if cond is not None else None,
                         shift.data_ptr() if shift is not None else None, out.data_ptr())
            return out
        # small batches: replay the whole forward (36 launches) from a CUDA graph over static buffers
        g = plan.graph
        if g is None or g[4] != (cond is not None) or g[2].shape != aux.shape:
            sx, sa = torch.empty_like(x), torch.empty_like(aux)
            so = torch.empty(B, 1, L, dtype=torch.float32, device=device)
            sx.copy_(x)
            sa.copy_(aux)
            args = (plan, sx.data_ptr(), sa.data_ptr() if cond is not None else None,
                    sa.data_ptr() if cond is None else None, so.data_ptr())
            self._launch(*args)                       # warm-up outside capture (module loading, attribute opt-ins)
            graph = torch.cuda.CUDAGraph()
            # (an explicit capture stream on THIS device: torch.cuda.graph's default capture stream is a process-wide
            #  singleton created on whichever device captured first)
            with torch.cuda.graph(graph, stream=torch.cuda.Stream(device=device)):
                self._launch(*args)
            g = plan.graph = (graph, sx, sa, so, cond is not None)
        g[1].copy_(x)
        g[2].copy_(aux)
        g[0].replay()
        return g[3].clone()

    @torch.no_grad()
    def forward(self, mixtures, conditions):
        self._check_inputs(mixtures)
        B, _, L = mixtures.shape
        device = mixtures.device
        with torch.cuda.device(device):               # plan creation sets per-device kernel attributes; launches go to this device
            plan = self._get_plan(B, L, device)
            x = mixtures.detach().to(torch.float32).contiguous()
            c = conditions.detach().to(device=device, dtype=torch.float32).contiguous()
            if c.shape != (B, self._packed["struct"].condition_size):
                raise ValueError("condition must be (batch, %d); got %s" % (self._packed["struct"].condition_size,
                                                                           tuple(c.shape)))
            return self._run(plan, x, c, None, device)

    @torch.no_grad()
    def forward_film_dict(self, mixtures, film_dict):
        """``base(mixtures=, film_dict=)``: betas supplied by the caller (reference models/resunet.py:685-688)."""
        self._check_inputs(mixtures)
        B, _, L = mixtures.shape
        device = mixtures.device
        if self.film is not None:
            raise RuntimeError("engine built with a FiLM module; use forward()")
        betas = []
        for name in _ENC:
            d = film_dict[name]["conv_block1"]
            betas += [d["beta1"], d["beta2"]]
        for name in _DEC:
            d = film_dict[name]
            betas += [d["beta1"], d["conv_block2"]["beta1"], d["conv_block2"]["beta2"]]
        with torch.cuda.device(device):
            plan = self._get_plan(B, L, device)
            packed = self._packed
            beta = torch.cat([b.reshape(b.shape[0], -1).to(device=device, dtype=torch.float32).expand(B, -1)
                              for b in betas], dim=1)
            shift = (beta + packed["film"][1][None, :]).contiguous()      # + folded BN shift
            x = mixtures.detach().to(torch.float32).contiguous()
            return self._run(plan, x, None, shift, device)

    @torch.no_grad()
    def forward_stages(self, mixtures, conditions, out, stage_mask):
        """Run a subset of the stages (bench.py brackets them with CUDA events). Inputs as in forward(): contiguous fp32
        CUDA tensors, module in eval mode."""
        self._check_inputs(mixtures)
        for t in (mixtures, conditions, out):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.device == mixtures.device):
                raise ValueError("forward_stages takes contiguous fp32 CUDA tensors on one device")
        B, _, L = mixtures.shape
        with torch.cuda.device(mixtures.device):
            plan = self._get_plan(B, L, mixtures.device)
            _cabi.check(_cabi.load().lass_resunet30_forward_stages(
                plan.handle, stage_mask, mixtures.data_ptr(), conditions.data_ptr(), None, out.data_ptr(),
                self.stft_precision_mode, torch.cuda.current_stream().cuda_stream))

    def unet_flops(self, B, L, device):
        return _cabi.load().lass_resunet30_unet_flops(self._get_plan(B, L, device).handle)

    def time_unet_launches(self, B, L, device):
        """Per-launch (ms, algorithmic FLOPs) of the UNET stage over the workspace's current contents (debug / profiling)."""
        plan = self._get_plan(B, L, device)
        ms = (ctypes.c_float * 64)()
        fl = (ctypes.c_double * 64)()
        with torch.cuda.device(device):
            n = _cabi.load().lass_debug_time_unet_launches(plan.handle, ms, fl, 64, torch.cuda.current_stream().cuda_stream)
        if n < 0:
            _cabi.check(n)
        return [(ms[i], fl[i]) for i in range(n)]

    def raw_stream_peak(self, B, L, device):
        """Largest |value| in the saturating-fp16 raw residual / skip tensors of the LAST forward with this (B, L) (run it with
        ``use_graphs = False``).  Values approaching 65504 mean a checkpoint's activations do not fit the fp16 raw stream
        (tests/test_gpu_forward.py::test_raw_stream_fp16_headroom): the output would clip silently."""
        names = ["x_raw%d" % k for k in range(1, 7)] + ["cat_raw%d" % k for k in range(6)]
        return max(float(self.debug_buffer(B, L, device, n).float().abs().max()) for n in names)

    def num_launches(self, B, L, device):
        return _cabi.load().lass_resunet30_num_launches(self._get_plan(B, L, device).handle)

    def debug_buffer(self, B, L, device, name):
        """Intermediate tensor of the last forward with this (B, L) — a COPY, for tests."""
        plan = self._get_plan(B, L, device)
        dims = (ctypes.c_int * 4)()
        eb = ctypes.c_int()
        ptr = _cabi.load().lass_resunet30_buffer(plan.handle, name.encode(), ctypes.byref(dims), ctypes.byref(eb))
        if not ptr:
            raise KeyError(name)
        off = ptr - plan.workspace.data_ptr()
        n = dims[0] * dims[1] * dims[2] * dims[3] * eb.value
        raw = plan.workspace[off:off + n].clone()
        if eb.value == 4:
            return raw.view(torch.float32).reshape(*dims)
        dt = torch.float16 if (name.startswith("x_raw") or name.startswith("cat_raw")) else torch.bfloat16
        return raw.view(dt).reshape(*dims)
