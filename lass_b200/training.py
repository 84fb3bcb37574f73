"""Training step of the separation model on the sm_100a kernels (BASELINE config 4).

Mirrors what the reference does per step and rank (``models/audiosep.py:52-145``, ``losses.py:4-9``, DDP at
``train.py:266-283``):

    ss_model.train(); out = ss_model(input_dict)['waveform']      batch-statistics BatchNorm, running stats updated
    loss = l1_wav(out, segment); loss.backward()                   gradients of all 26.4 M live parameters
    DDP gradient all-reduce (NCCL)                                 the path's only collective
    AdamW(amsgrad=True).step(); LambdaLR.step()

``TrainEngine`` runs the forward and the hand-derived backward as a fixed sequence of kernel launches over buffers it
owns (NHWC, raw conv outputs fp16, activated tensors and gradients bf16, statistics / parameter gradients fp32):

* forward / dgrad convolutions: the tcgen05 implicit-GEMM kernel of the inference path (``lass_conv_*``), dgrad with
  flipped / transposed weights; transposed-conv dgrad as a 1x1 conv over the un-shuffled gradient;
* wgrad: ``lass_wgrad`` (pixels are the GEMM K dimension);
* BatchNorm batch statistics, activation, backward reductions: memory-bound kernels (``lass_bn_*``);
* spectral ends: K1 / K5 forward, ``lass_istft_bwd`` (the adjoint of the ISTFT is an STFT of the window-sum-normalised
  gradient) + ``lass_mask_bwd``;
* parameters, gradients and optimizer state live in flat fp32 buffers: ONE (three-bucket) NCCL all-reduce, one fused
  AdamW-amsgrad launch, then the 16-bit kernel layouts of the conv weights are refreshed.

All arithmetic is in ``kernels`` (default: ``lass_b200.train_kernels``, the C ABI; no CPU fallback).
"""
import contextlib
import functools

import torch

from .engine import _DEC, _ENC


def _on_device(method):
    """Run a TrainEngine method with the engine's device current: the kernels launch on torch's current stream of the
    CURRENT device, and kernel attributes / SM counts are per device."""
    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        guard = torch.cuda.device(self.device) if self.device.type == "cuda" else contextlib.nullcontext()
        with guard:
            return method(self, *args, **kwargs)
    return wrapper

ENC = ((32, 32, (2, 2)), (32, 64, (2, 2)), (64, 128, (2, 2)), (128, 256, (2, 2)), (256, 384, (2, 2)),
       (384, 384, (1, 2)), (384, 384, (1, 1)))                      # cin, cout, pool   (models/resunet.py:315-370)
DEC = ((384, 384, (1, 2)), (384, 384, (2, 2)), (384, 256, (2, 2)), (256, 128, (2, 2)), (128, 64, (2, 2)),
       (64, 32, (2, 2)))                                            # cin, cout, upsample (models/resunet.py:371-418)
BN_MOMENTUM = 0.01
BN_EPS = 1e-5
# all-reduce bucket B1 = encoder blocks 6 .. _B1_LAST_LEVEL (10.8 of the encoder's 11.7 M parameters): reduced while the high-resolution
# blocks' backward (most of the encoder's time) still runs
_B1_LAST_LEVEL = 4


def film_row_offsets():
    """First row of every FiLM / activation site in the (B, J) beta table — the order of include/lass_b200.h
    (site = 2k + {0,1} encoder, 14 + 3j + {0,1,2} decoder)."""
    off, o = [], 0
    for cin, cout, _ in ENC:
        off += [o, o + cin]
        o += cin + cout
    for cin, cout, _ in DEC:
        off += [o, o + cin, o + cin + 2 * cout]
        o += cin + 3 * cout
    return off, o


class _Site:
    """One executed BatchNorm (+ FiLM beta) site."""

    def __init__(self, bn, row, C):
        self.bn, self.row, self.C = bn, row, C


class TrainEngine:
    def __init__(self, model, kernels=None, sync_batchnorm=False, process_group=None, sync_transport="nccl"):
        """``sync_batchnorm=True`` = the reference's training configuration (``sync_batchnorm: True``,
        config/audiosep_base.yaml:42 -> ``torch.nn.SyncBatchNorm`` under DDP, train.py:176,266-283): every BatchNorm's batch
        statistics (and the two sums of its backward) are taken over the clips of ALL ranks of ``process_group`` -- one small
        all-reduce per BatchNorm site in the forward and one in the backward.  Without torch.distributed (or world size 1)
        it changes nothing.  Default False: statistics per rank (what plain DDP does).  ``sync_transport="p2p"`` (CUDA, one
        node): the exchange is fused into the finalize kernels over NVLink peer memory (symmetric memory; one launch per site
        and direction, no collective call) instead of one NCCL all-reduce per site (``"nccl"``)."""
        if kernels is None:
            from . import train_kernels as kernels
        self.k = kernels
        self.model = model
        self.sync_batchnorm = bool(sync_batchnorm)
        self.process_group = process_group
        if sync_transport not in ("nccl", "p2p"):
            raise ValueError("sync_transport: 'nccl' or 'p2p'")
        self.sync_transport = sync_transport
        self._p2p = None              # forward-side symmetric buffers (created collectively on first use)
        self._epoch = 0               # forward calls so far: the flag value peers wait for, its parity picks the sums half
        base, film = model.base, model.film
        if base.input_channels != 1 or base.output_channels != 1:
            raise NotImplementedError("training is implemented for input_channels == output_channels == 1")
        self.n_fft, self.hop = base.window_size, base.hop_size
        self.device = base.pre_conv.weight.device
        self.rows, self.J = film_row_offsets()
        self.K = film.condition_size
        self._flatten_parameters()
        self._make_sites()
        self._alloc_weights()
        self.refresh_weights()
        self._ws = {}
        self.step_count = 0

    # ------------------------------------------------------------------ flat parameter / gradient buffers
    def _flatten_parameters(self):
        base, film = self.model.base, self.model.film
        encb = [getattr(base, n).conv_block1 for n in _ENC]
        decb = [getattr(base, n) for n in _DEC]
        entries = []

        def add(name, p):
            entries.append((name, p))

        def add_block(prefix, cb):
            for n in ("bn1", "bn2"):
                add(prefix + n + ".weight", getattr(cb, n).weight)
                add(prefix + n + ".bias", getattr(cb, n).bias)
            add(prefix + "conv1.weight", cb.conv1.weight)
            add(prefix + "conv2.weight", cb.conv2.weight)
            if cb.is_shortcut:
                add(prefix + "shortcut.weight", cb.shortcut.weight)
                add(prefix + "shortcut.bias", cb.shortcut.bias)

        # bucket A: everything whose gradient is final once the decoder's backward is (all-reduced while the encoder's runs)
        add("after.w", base.after_conv.weight)
        add("after.b", base.after_conv.bias)
        for j in reversed(range(6)):
            blk = decb[j]
            add_block("dec%d.cb2." % j, blk.conv_block2)
            add("dec%d.up" % j, blk.conv1.weight)
            add("dec%d.bn1.weight" % j, blk.bn1.weight)
            add("dec%d.bn1.bias" % j, blk.bn1.bias)
        n_a = len(entries)
        # bucket B1: the deep encoder blocks (most of the encoder's bytes, final early in its backward); B2: the rest, pre_conv, bn0, FiLM
        n_b1 = None
        for k in reversed(range(7)):
            if k == _B1_LAST_LEVEL - 1:
                n_b1 = len(entries)
            add_block("enc%d." % k, encb[k])
        add("pre.w", base.pre_conv.weight)
        add("pre.b", base.pre_conv.bias)
        add("bn0.weight", base.bn0.weight)
        add("bn0.bias", base.bn0.bias)
        from .engine import film_sites
        self._film_names = [fname for _bn, fname in film_sites(base)]
        film_w_first = len(entries)
        for fname in self._film_names:
            add("film.w." + fname, getattr(film, fname).weight)
        film_b_first = len(entries)
        for fname in self._film_names:
            add("film.b." + fname, getattr(film, fname).bias)
        n_live = len(entries)
        # dead parameters (never receive a gradient: reference decoder_blockN.bn2 and film decoder_blockN->beta2)
        seen = {id(p) for _n, p in entries}
        for name, p in self.model.named_parameters():
            if p.requires_grad and id(p) not in seen:
                add("dead." + name, p)
        offs, o = [], 0
        for _name, p in entries:
            offs.append(o)
            o += (p.numel() + 31) // 32 * 32
        total = o
        self.P = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.G = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.index = {}
        self.params = []
        with torch.no_grad():
            for (name, p), off in zip(entries, offs):
                view = self.P[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view                                        # the module's parameters now ARE views of the flat buffer
                self.index[name] = (off, p)
                self.params.append(p)
        self.bucket_a_end = offs[n_a]
        self.bucket_b1_end = offs[n_b1]
        self.live_end = offs[n_live] if n_live < len(entries) else total
        self.film_w_off, self.film_b_off = offs[film_w_first], offs[film_b_first]
        assert offs[film_b_first] - offs[film_w_first] == self.J * self.K, "FiLM weights must be contiguous"
        self.opt_state = None

    def g(self, name):
        off, p = self.index[name]
        return self.G[off:off + p.numel()].view(p.shape)

    def p(self, name):
        return self.index[name][1]

    def _make_sites(self):
        base = self.model.base
        self.site = {}
        for k, name in enumerate(_ENC):
            cb = getattr(base, name).conv_block1
            self.site[2 * k] = _Site(cb.bn1, self.rows[2 * k], ENC[k][0])
            self.site[2 * k + 1] = _Site(cb.bn2, self.rows[2 * k + 1], ENC[k][1])
        for j, name in enumerate(_DEC):
            blk = getattr(base, name)
            self.site[14 + 3 * j] = _Site(blk.bn1, self.rows[14 + 3 * j], DEC[j][0])
            self.site[14 + 3 * j + 1] = _Site(blk.conv_block2.bn1, self.rows[14 + 3 * j + 1], 2 * DEC[j][1])
            self.site[14 + 3 * j + 2] = _Site(blk.conv_block2.bn2, self.rows[14 + 3 * j + 2], DEC[j][1])
        names = {}
        for k in range(7):
            names[2 * k], names[2 * k + 1] = "enc%d.bn1" % k, "enc%d.bn2" % k
        for j in range(6):
            names[14 + 3 * j], names[14 + 3 * j + 1], names[14 + 3 * j + 2] = \
                "dec%d.bn1" % j, "dec%d.cb2.bn1" % j, "dec%d.cb2.bn2" % j
        self.site_name = names
        dev = self.device
        # One flat buffer each for the forward sums and num_batches_tracked: cleared / incremented ONCE per step, not per site.
        order = sorted(self.site)
        self._sums_flat = torch.zeros(sum(2 * self.site[s].C for s in order), dtype=torch.float64, device=dev)
        self._nbt = torch.zeros(len(order) + 1, dtype=torch.int64, device=dev)
        o = 0
        for i, s in enumerate(order):
            st = self.site[s]
            st.bnp = torch.zeros(6 * st.C, dtype=torch.float32, device=dev)
            st.sums = self._sums_flat[o:o + 2 * st.C].view(2, st.C)
            o += 2 * st.C
            self._adopt_nbt(st.bn, i)
        self._adopt_nbt(base.bn0, len(order))
        F = self.n_fft // 2 + 1
        self.bnp0 = torch.zeros(6 * F, dtype=torch.float32, device=dev)
        self.sums0 = torch.zeros(2, F, dtype=torch.float64, device=dev)

    def _adopt_nbt(self, bn, i):
        """bn.num_batches_tracked becomes a view of the flat counter buffer (one `+= 1` per step for all 34 BatchNorms)."""
        with torch.no_grad():
            self._nbt[i] = bn.num_batches_tracked.to(self._nbt.device)
            bn.num_batches_tracked.data = self._nbt[i]

    # ------------------------------------------------------------------ 16-bit kernel layouts of the conv weights
    def _alloc_weights(self):
        k, dev = self.k, self.device
        base = self.model.base
        self.w = {}

        def conv_pair(name, conv, fwd_dtype):
            co, ci, kh, kw = conv.weight.shape
            taps = kh * kw
            self.w[name] = (conv.weight, k.KIND_CONV, k.empty((taps, co, ci), fwd_dtype, dev),
                            k.empty((taps, ci, co), k.GRAD_DTYPE, dev))

        for kk, name in enumerate(_ENC):
            cb = getattr(base, name).conv_block1
            conv_pair("enc%d.conv1" % kk, cb.conv1, k.ACT_DTYPE)
            conv_pair("enc%d.conv2" % kk, cb.conv2, k.ACT_DTYPE)
            if cb.is_shortcut:
                conv_pair("enc%d.sc" % kk, cb.shortcut, k.RAW_DTYPE)
            else:
                c = ENC[kk][0]
                self.w["enc%d.sc" % kk] = (None, None, torch.eye(c, dtype=k.RAW_DTYPE, device=dev).reshape(1, c, c), None)
        for j, name in enumerate(_DEC):
            blk = getattr(base, name)
            cin, cout, (uh, uw) = DEC[j]
            self.w["dec%d.up" % j] = (blk.conv1.weight, k.KIND_CONVT, k.empty((1, uh * uw * cout, cin), k.ACT_DTYPE, dev),
                                     k.empty((1, cin, uh * uw * cout), k.GRAD_DTYPE, dev))
            cb = blk.conv_block2
            conv_pair("dec%d.conv1" % j, cb.conv1, k.ACT_DTYPE)
            conv_pair("dec%d.conv2" % j, cb.conv2, k.ACT_DTYPE)
            conv_pair("dec%d.sc" % j, cb.shortcut, k.RAW_DTYPE)

        # multi-tensor tables: every conv weight is re-packed by ONE launch after the optimizer step; the weight gradients the
        # tcgen05 kernel leaves in its packed (taps, co, ci) layout (in Gp, at the parameter's offset) are un-packed into G by one
        # launch per all-reduce bucket
        self._pack_table = k.pack_weights_table([(param.data, kind, fwd, dgrad) for (param, kind, fwd, dgrad) in self.w.values()
                                                 if param is not None], dev)
        self.Gp = torch.zeros_like(self.G)
        items = ([], [], [])
        for name, (off, p) in self.index.items():
            if p.dim() != 4 or name.startswith(("dead.", "after.", "pre.")):
                continue
            kind = k.KIND_CONVT if name.endswith(".up") else k.KIND_CONV
            if kind == k.KIND_CONV and p.shape[2] * p.shape[3] == 1:
                continue                                           # (co, ci, 1, 1) IS the packed layout: wgrad writes G directly
            n = p.numel()
            bucket = 0 if off < self.bucket_a_end else (1 if off < self.bucket_b1_end else 2)
            items[bucket].append((self.Gp[off:off + n], kind, self.G[off:off + n].view(p.shape)))
        self._unpack_a, self._unpack_b1, self._unpack_b2 = (k.unpack_grads_table(it, dev) for it in items)

    @_on_device
    def refresh_weights(self):
        """Re-derive the 16-bit kernel layouts from the fp32 parameters (after an optimizer step / load_state_dict)."""
        self.k.pack_weights(self._pack_table)

    # ------------------------------------------------------------------ per-(B, L) workspace
    def _workspace(self, B, L):
        key = (B, L)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        k, dev = self.k, self.device
        n_fft, hop = self.n_fft, self.hop
        T, F = L // hop + 1, n_fft // 2 + 1
        Tp, Fp = (T + 31) // 32 * 32, n_fft // 2
        H, W = [Tp], [Fp]
        for _ in range(5):
            H.append(H[-1] // 2)
            W.append(W[-1] // 2)
        H.append(H[5])
        W.append(W[5] // 2)

        class WS:
            pass

        ws = WS()
        ws.B, ws.L, ws.T, ws.F, ws.Tp, ws.Fp, ws.H, ws.W = B, L, T, F, Tp, Fp, H, W
        R, A, G = k.RAW_DTYPE, k.ACT_DTYPE, k.GRAD_DTYPE

        def buf(lvl, c, dt):
            return k.empty((B, H[lvl], W[lvl], c), dt, dev)

        ws.beta = torch.zeros(B, self.J, dtype=torch.float32, device=dev)
        ws.dbeta = torch.zeros(B, self.J, dtype=torch.float32, device=dev)
        ws.x_raw = [buf(kk, ENC[kk][0], R) for kk in range(7)]
        ws.x_act = [buf(kk, ENC[kk][0], A) for kk in range(7)]
        ws.h_raw = [buf(kk, ENC[kk][1], R) for kk in range(7)]
        ws.a2 = [buf(kk, ENC[kk][1], A) for kk in range(7)]
        ws.cat_raw = [buf(kk, 2 * ENC[kk][1], R) for kk in range(6)]
        ws.cat_act = [buf(kk, 2 * ENC[kk][1], A) for kk in range(6)]
        ws.hd_raw = [buf(kk, ENC[kk][1], R) for kk in range(6)]
        ws.a2d = [buf(kk, ENC[kk][1], A) for kk in range(6)]
        ws.d_raw = [buf(kk, ENC[kk][1], R) for kk in range(7)]      # [6] = conv_block7a output, [5..0] = decoder block outputs
        ws.xin_act = [None] + [buf(6 - j, DEC[j][0], A) for j in reversed(range(6))]     # indexed by the level it lives on
        ws.feat = torch.zeros(B, 3, Tp, Fp, dtype=torch.float32, device=dev)
        ws.bsums_flat = torch.zeros(sum(B * st.C * 2 for st in self.site.values()), dtype=torch.float32, device=dev)
        ws.bsums, o = {}, 0
        for s, st in self.site.items():
            ws.bsums[s] = ws.bsums_flat[o:o + B * st.C * 2].view(B, st.C, 2)
            o += B * st.C * 2
        # SyncBatchNorm backward: per-channel totals (C, 2) fp64 of every site, the all-reduce payloads
        ws.btotals = {s: torch.zeros(st.C, 2, dtype=torch.float64, device=dev) for s, st in self.site.items()}
        ws.sync_world = 1
        # gradients (bf16)
        ws.g_y = [buf(kk, ENC[kk][1], G) for kk in range(7)]
        ws.g_a2 = [buf(kk, ENC[kk][1], G) for kk in range(7)]
        ws.g_h = [buf(kk, ENC[kk][1], G) for kk in range(7)]
        ws.g_catact = [buf(kk, 2 * ENC[kk][1], G) for kk in range(6)]
        ws.g_sc_d = [buf(kk, 2 * ENC[kk][1], G) for kk in range(6)]
        ws.g_cat = [buf(kk, 2 * ENC[kk][1], G) for kk in range(6)]
        ws.g_xact = [buf(kk, ENC[kk][0], G) for kk in range(7)]
        ws.g_sc_e = [buf(kk, ENC[kk][0], G) if ENC[kk][0] != ENC[kk][1] else None for kk in range(7)]
        ws.g_xraw = [buf(kk, ENC[kk][0], G) for kk in range(7)]
        ws.dU = [None] * 7
        ws.g_xinact = [None] * 7
        for j in range(6):
            cin, cout, (uh, uw) = DEC[j]
            ws.dU[6 - j] = buf(6 - j, uh * uw * cout, G)
            ws.g_xinact[6 - j] = buf(6 - j, cin, G)
        ws.dfeat = torch.zeros(B, 3, Tp, Fp, dtype=torch.float32, device=dev)
        ws.dre = torch.zeros(B, T, F, dtype=torch.float32, device=dev)
        ws.dim = torch.zeros(B, T, F, dtype=torch.float32, device=dev)
        ws.dwave = torch.zeros(B, L, dtype=torch.float32, device=dev)
        ws.loss_sum = torch.zeros(1, dtype=torch.float32, device=dev)
        ws.stft_ws = self._stft_workspace(B, L)
        self._build_convs(ws)
        if len(self._ws) >= 2:
            self._ws.pop(next(iter(self._ws)))
        self._ws[key] = ws
        return ws

    def _stft_workspace(self, B, L):
        if self.device.type != "cuda":
            return None
        from . import _cabi
        need = _cabi.load().lass_stft_workspace_bytes(B, L, self.n_fft, self.hop)
        return torch.empty(need + 256, dtype=torch.uint8, device=self.device)

    def _spectral_tables(self):
        t = getattr(self, "_spec", None)
        if t is None:
            from . import packing
            base = self.model.base
            window, tw = packing.istft_tables(self.n_fft, device=self.device)
            if self.device.type == "cuda":
                hi, lo = packing.pack_stft_basis(base.stft.conv_real.weight.data, base.stft.conv_imag.weight.data)
            else:       # the CPU emulation of the kernel interface takes the reference's conv weights directly
                hi, lo = base.stft.conv_real.weight.data, base.stft.conv_imag.weight.data
            t = self._spec = (hi, lo, window, tw)
        return t

    def _build_convs(self, ws):
        C = self.k.ConvSpec
        B, H, W = ws.B, ws.H, ws.W
        base = self.model.base
        cv = ws.conv = {}
        for kk in range(7):
            cin, cout, pool = ENC[kk]
            w1, w2, wsc = self.w["enc%d.conv1" % kk], self.w["enc%d.conv2" % kk], self.w["enc%d.sc" % kk]
            cb = getattr(base, _ENC[kk]).conv_block1
            cv["enc%d.c1" % kk] = C(B, H[kk], W[kk], cout, [(ws.x_act[kk], 0, cin, w1[2], 9)], full_raw=ws.h_raw[kk])
            segs = [(ws.a2[kk], 0, cout, w2[2], 9), (ws.x_raw[kk], 0, cin, wsc[2], 1)]
            bias = cb.shortcut.bias if cb.is_shortcut else None
            if kk < 6:
                cv["enc%d.c2" % kk] = C(B, H[kk], W[kk], cout, segs, bias=bias, full_raw=ws.cat_raw[kk], full_raw_coff=cout,
                                        pool=(pool[0], 2), pool_raw=ws.x_raw[kk + 1])
            else:
                cv["enc%d.c2" % kk] = C(B, H[kk], W[kk], cout, segs, bias=bias, full_raw=ws.d_raw[6])
            # backward
            cv["enc%d.c2.dgrad" % kk] = C(B, H[kk], W[kk], cout, [(ws.g_y[kk], 0, cout, w2[3], 9)], full_raw=ws.g_a2[kk])
            cv["enc%d.c1.dgrad" % kk] = C(B, H[kk], W[kk], cin, [(ws.g_h[kk], 0, cout, w1[3], 9)], full_raw=ws.g_xact[kk])
            if cb.is_shortcut:
                cv["enc%d.sc.dgrad" % kk] = C(B, H[kk], W[kk], cin, [(ws.g_y[kk], 0, cout, wsc[3], 1)], full_raw=ws.g_sc_e[kk])
        for j in range(6):
            cin, cout, (uh, uw) = DEC[j]
            lin, lo = 6 - j, 5 - j
            wu, w1, w2, wsc = (self.w["dec%d.%s" % (j, n)] for n in ("up", "conv1", "conv2", "sc"))
            blk = getattr(base, _DEC[j])
            cv["dec%d.up" % j] = C(B, H[lin], W[lin], uh * uw * cout, [(ws.xin_act[lin], 0, cin, wu[2], 1)], up=(uh, uw),
                                   full_raw=ws.cat_raw[lo], full_raw_coff=0)
            cv["dec%d.c1" % j] = C(B, H[lo], W[lo], cout, [(ws.cat_act[lo], 0, 2 * cout, w1[2], 9)], full_raw=ws.hd_raw[lo])
            after = None
            if j == 5:
                self._after_w = base.after_conv.weight.data.view(3, 32)
                after = (self._after_w, base.after_conv.bias.data, ws.feat)
            cv["dec%d.c2" % j] = C(B, H[lo], W[lo], cout, [(ws.a2d[lo], 0, cout, w2[2], 9), (ws.cat_raw[lo], 0, 2 * cout, wsc[2], 1)],
                                   bias=blk.conv_block2.shortcut.bias, full_raw=ws.d_raw[lo], after=after)
            cv["dec%d.c2.dgrad" % j] = C(B, H[lo], W[lo], cout, [(ws.g_y[lo], 0, cout, w2[3], 9)], full_raw=ws.g_a2[lo])
            cv["dec%d.sc.dgrad" % j] = C(B, H[lo], W[lo], 2 * cout, [(ws.g_y[lo], 0, cout, wsc[3], 1)], full_raw=ws.g_sc_d[lo])
            cv["dec%d.c1.dgrad" % j] = C(B, H[lo], W[lo], 2 * cout, [(ws.g_h[lo], 0, cout, w1[3], 9)], full_raw=ws.g_catact[lo])
            cv["dec%d.up.dgrad" % j] = C(B, H[lin], W[lin], cin, [(ws.dU[lin], 0, uh * uw * cout, wu[3], 1)],
                                         full_raw=ws.g_xinact[lin])

    # ------------------------------------------------------------------ SyncBatchNorm plumbing
    def _sync_world(self):
        """Number of ranks whose clips share BatchNorm statistics (1 = statistics per rank)."""
        if not self.sync_batchnorm:
            return 1
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return 1
        return dist.get_world_size(self.process_group)

    def _collective_stream(self):
        """The ONE stream every collective of this engine is issued on (gradient buckets and BatchNorm statistics alike): NCCL
        operations of one communicator must not run concurrently from two streams -- a bucket all-reduce in flight on a side
        stream while a statistics all-reduce starts on the compute stream deadlocks (seen on 2 x B200)."""
        comm = getattr(self, "_comm_stream", None)
        if comm is None:
            comm = self._comm_stream = torch.cuda.Stream(device=self.device)
        return comm

    def _all_reduce_sums(self, t):
        """Sum a small fp64 statistics tensor over the ranks, in stream order with the kernels around it."""
        import torch.distributed as dist
        if self.device.type != "cuda":
            dist.all_reduce(t, group=self.process_group)
            return
        comm, cur = self._collective_stream(), torch.cuda.current_stream()
        comm.wait_stream(cur)
        with torch.cuda.stream(comm):
            dist.all_reduce(t, group=self.process_group)
        cur.wait_stream(comm)

    # flag_index of the peer exchange: forward site s -> s, bn0 -> 32, backward site s -> 64 + s
    _FLAG_ROWS = 128

    def _p2p_active(self, ws):
        return ws.sync_world > 1 and self.sync_transport == "p2p" and self.device.type == "cuda"

    def _p2p_group(self):
        import torch.distributed as dist
        return self.process_group if self.process_group is not None else dist.group.WORLD

    def _p2p_setup(self):
        """COLLECTIVE (every rank, same point of the program): the forward sums of all sites (+ bn0), twice (epoch parity), and
        the flag table, in symmetric memory mapped by every rank of the node."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        from types import SimpleNamespace
        group, dev = self._p2p_group(), self.device
        order = sorted(self.site)
        F = self.n_fft // 2 + 1
        total = sum(2 * self.site[s].C for s in order) + 2 * F
        buf = symm.empty(2 * total, dtype=torch.float64, device=dev)
        flags = symm.empty(self._FLAG_ROWS * 16, dtype=torch.int64, device=dev)
        buf.zero_()
        flags.zero_()
        torch.cuda.synchronize(dev)
        h_buf, h_fl = symm.rendezvous(buf, group), symm.rendezvous(flags, group)
        dist.barrier(group)                     # every rank's flags are zero before anyone publishes an epoch
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        p = SimpleNamespace(buf=buf, flags=flags, handles=(h_buf, h_fl), total=total, status=status, offsets={}, views=[],
                            flag_ptrs=list(h_fl.buffer_ptrs), rank=dist.get_rank(group),
                            table=self.k.PeerTable(h_buf.buffer_ptrs, h_fl.buffer_ptrs, dist.get_rank(group), status))
        for h in range(2):
            flat = buf[h * total:(h + 1) * total]
            sites, o = {}, 0
            for s_ in order:
                C = self.site[s_].C
                sites[s_] = flat[o:o + 2 * C].view(2, C)
                p.offsets[s_] = o
                o += 2 * C
            p.offsets["bn0"] = o
            p.views.append(SimpleNamespace(flat=flat, sites=sites, sums0=flat[o:o + 2 * F].view(2, F)))
        self._p2p = p

    def _p2p_setup_ws(self, ws):
        """COLLECTIVE: this workspace's backward sums (B, C, 2) of all sites, twice, in symmetric memory."""
        import torch.distributed._symmetric_memory as symm
        from types import SimpleNamespace
        n_sums = sum(ws.B * st.C * 2 for st in self.site.values())          # per-clip sums, fp32 (even: 8-byte aligned end)
        total = n_sums + sum(st.C * 4 for st in self.site.values())          # + per-channel totals (C, 2) fp64 = 4 floats per channel
        buf = symm.empty(2 * total, dtype=torch.float32, device=self.device)
        buf.zero_()
        torch.cuda.synchronize(self.device)
        h_buf = symm.rendezvous(buf, self._p2p_group())
        p = SimpleNamespace(buf=buf, handle=h_buf, total=total, offsets={}, tot_offsets={}, views=[],
                            table=self.k.PeerTable(h_buf.buffer_ptrs, self._p2p.flag_ptrs, self._p2p.rank, self._p2p.status))
        for h in range(2):
            flat = buf[h * total:(h + 1) * total]
            sites, o = {}, 0
            for s_, st in self.site.items():
                sites[s_] = flat[o:o + ws.B * st.C * 2].view(ws.B, st.C, 2)
                p.offsets[s_] = o
                o += ws.B * st.C * 2
            p.views.append(SimpleNamespace(flat=flat[:n_sums], sites=sites))
        o = n_sums
        for s_, st in self.site.items():                                     # float offsets of the totals areas inside a half
            p.tot_offsets[s_] = o
            o += st.C * 4
        ws.p2p = p

    def check_sync_status(self):
        """Raise if a peer-memory exchange ever timed out (a rank that never arrived); synchronises the device."""
        if self._p2p is not None and int(self._p2p.status.item()) != 0:
            raise RuntimeError("sync_batchnorm over peer memory: a rank never published its statistics (spin limit reached)")

    # ------------------------------------------------------------------ forward (train mode)
    def _bn_fwd(self, ws, site, x, x_coff, out, out_coff):
        """Batch statistics of x[..., x_coff:+C] -> scale / shift, running-stat update, out = lrelu(bn(x) + beta)."""
        k, st = self.k, self.site[site]
        bn = st.bn
        count = x.shape[0] * x.shape[1] * x.shape[2]
        if self._p2p_active(ws):              # statistics exchange fused into the finalize kernel (NVLink peer memory)
            p = self._p2p
            h = ws.epoch & 1
            k.bn_stats_acc(x, x_coff, st.C, p.views[h].sites[site])
            k.bn_finalize_p2p(p.table, h * p.total + p.offsets[site], site, ws.epoch, count * ws.sync_world, bn.weight.data,
                              bn.bias.data, bn.running_mean, bn.running_var, BN_MOMENTUM, BN_EPS, st.bnp)
            k.bn_act(x, x_coff, out, out_coff, st.C, st.bnp, ws.beta[:, st.row:st.row + st.C])
            return
        k.bn_stats_acc(x, x_coff, st.C, st.sums)
        if ws.sync_world > 1:                 # torch.nn.SyncBatchNorm: mean / variance over every rank's pixels
            self._all_reduce_sums(st.sums)
            count *= ws.sync_world
        k.bn_finalize(st.sums, count, bn.weight.data, bn.bias.data, bn.running_mean, bn.running_var, BN_MOMENTUM, BN_EPS,
                      st.bnp)
        k.bn_act(x, x_coff, out, out_coff, st.C, st.bnp, ws.beta[:, st.row:st.row + st.C])

    @_on_device
    def forward(self, mixture, condition):
        """mixture (B, 1, L), condition (B, K) -> waveform (B, 1, L); keeps everything the backward needs."""
        k = self.k
        B, _, L = mixture.shape
        ws = self._workspace(B, L)
        base = self.model.base
        hi, lo, window, tw = self._spectral_tables()
        ws.sync_world = self._sync_world()
        self._epoch += 1
        ws.epoch = self._epoch
        if self._p2p_active(ws):
            if self._p2p is None:
                self._p2p_setup()
            if getattr(ws, "p2p", None) is None:
                self._p2p_setup_ws(ws)
            self._p2p.views[ws.epoch & 1].flat.zero_()
        else:
            self._sums_flat.zero_()      # bn_stats_acc adds into the sums of all sites: one memset per step
        self._nbt += 1                   # every BatchNorm's num_batches_tracked (views of this buffer)
        ws.cond = condition.detach().to(torch.float32).contiguous()
        wave_in = mixture.detach().to(torch.float32).reshape(B, L).contiguous()
        ws.mag, ws.cos, ws.sin = k.stft(wave_in, hi, lo, self.n_fft, self.hop, ws.stft_ws)
        # FiLM betas for every site: one GEMM over the contiguous slice of the flat parameter buffer
        film_w = self.P[self.film_w_off:self.film_w_off + self.J * self.K].view(self.J, self.K)
        film_b = self.P[self.film_b_off:self.film_b_off + self.J]
        k.film(ws.cond, film_w, film_b, ws.beta)
        # bn0 (per frequency bin over batch x time) + zero time padding + Nyquist drop + pre_conv
        bn0 = base.bn0
        if self._p2p_active(ws):
            p = self._p2p
            h = ws.epoch & 1
            k.bn0_stats(ws.mag, p.views[h].sums0)
            k.bn_finalize_p2p(p.table, h * p.total + p.offsets["bn0"], 32, ws.epoch, B * ws.T * ws.sync_world, bn0.weight.data,
                              bn0.bias.data, bn0.running_mean, bn0.running_var, BN_MOMENTUM, BN_EPS, self.bnp0)
        else:
            k.bn0_stats(ws.mag, self.sums0)
            if ws.sync_world > 1:
                self._all_reduce_sums(self.sums0)
            k.bn_finalize(self.sums0, B * ws.T * ws.sync_world, bn0.weight.data, bn0.bias.data, bn0.running_mean,
                          bn0.running_var, BN_MOMENTUM, BN_EPS, self.bnp0)
        k.pre_fwd(ws.mag, self.bnp0, base.pre_conv.weight.data.view(32), base.pre_conv.bias.data, ws.x_raw[0])
        cv = ws.conv
        for kk in range(7):
            self._bn_fwd(ws, 2 * kk, ws.x_raw[kk], 0, ws.x_act[kk], 0)
            k.conv(cv["enc%d.c1" % kk])
            self._bn_fwd(ws, 2 * kk + 1, ws.h_raw[kk], 0, ws.a2[kk], 0)
            k.conv(cv["enc%d.c2" % kk])
        for j in range(6):
            lin, lo_ = 6 - j, 5 - j
            cout = DEC[j][1]
            self._bn_fwd(ws, 14 + 3 * j, ws.d_raw[lin], 0, ws.xin_act[lin], 0)
            k.conv(cv["dec%d.up" % j])
            self._bn_fwd(ws, 14 + 3 * j + 1, ws.cat_raw[lo_], 0, ws.cat_act[lo_], 0)
            k.conv(cv["dec%d.c1" % j])
            self._bn_fwd(ws, 14 + 3 * j + 2, ws.hd_raw[lo_], 0, ws.a2d[lo_], 0)
            k.conv(cv["dec%d.c2" % j])
        wave = k.mask_istft(ws.feat, ws.mag, ws.cos, ws.sin, window, tw, self.n_fft, self.hop, L)
        ws.wave = wave
        self._last = ws
        return wave.view(B, 1, L)

    # ------------------------------------------------------------------ backward
    def _bn_bwd(self, ws, site, dact, x, x_coff, add, add_coff, dx, dx_coff):
        k, st = self.k, self.site[site]
        name = self.site_name[site]
        beta = ws.beta[:, st.row:st.row + st.C]
        count = x.shape[0] * x.shape[1] * x.shape[2]
        if self._p2p_active(ws):
            p = ws.p2p
            h = ws.epoch & 1
            k.bn_bwd_reduce_acc(dact, x, x_coff, st.C, st.bnp, beta, p.views[h].sites[site])
            k.bn_bwd_finalize_p2p(p.table, h * p.total + p.offsets[site], (h * p.total + p.tot_offsets[site]) // 2, 64 + site,
                                  ws.epoch, ws.B, count * ws.sync_world,
                                  st.bn.weight.data, st.bnp, self.g(name + ".weight"), self.g(name + ".bias"),
                                  ws.dbeta[:, st.row:st.row + st.C])
            k.bn_bwd_apply(dact, x, x_coff, st.C, st.bnp, beta, add, add_coff, dx, dx_coff)
            return
        sums = ws.bsums[site]
        k.bn_bwd_reduce_acc(dact, x, x_coff, st.C, st.bnp, beta, sums)
        if ws.sync_world > 1:
            # SyncBatchNorm backward: the input gradient takes the two sums over ALL ranks, the parameter gradients stay local
            # (the gradient all-reduce averages them like every other parameter's)
            totals = ws.btotals[site]
            k.bn_bwd_totals(sums, totals)
            self._all_reduce_sums(totals)
            k.bn_bwd_finalize_sync(sums, count * ws.sync_world, totals, st.bn.weight.data, st.bnp, self.g(name + ".weight"),
                                   self.g(name + ".bias"), ws.dbeta[:, st.row:st.row + st.C])
        else:
            k.bn_bwd_finalize(sums, count, st.bn.weight.data, st.bnp, self.g(name + ".weight"), self.g(name + ".bias"),
                              ws.dbeta[:, st.row:st.row + st.C])
        k.bn_bwd_apply(dact, x, x_coff, st.C, st.bnp, beta, add, add_coff, dx, dx_coff)

    def _wgrad(self, ws, name, kind, dy, co, x, ci, taps):
        k = self.k
        off, p = self.index[name]
        if taps == 1 and kind == k.KIND_CONV:
            k.wgrad(dy, 0, co, x, 0, ci, 1, self.G[off:off + p.numel()], acc=True)        # (co, ci, 1, 1) is the packed layout
            return
        k.wgrad(dy, 0, co, x, 0, ci, taps, self.Gp[off:off + p.numel()], acc=True)       # packed layout; un-packed per bucket (backward)

    def _block_bwd(self, ws, pfx, sites, dy, x_raw, x_act, h_raw, a2, g_a2, g_h, g_xact, g_sc, g_x, cin, cout, has_sc,
                   conv_pfx):
        """Backward of one ConvBlockRes (reference models/resunet.py:147-165).  dy: gradient of the block output."""
        k, cv = self.k, ws.conv
        if has_sc:
            k.channel_sum(dy, 0, cout, self.g(pfx + "shortcut.bias"), acc=True)
            self._wgrad(ws, pfx + "shortcut.weight", k.KIND_CONV, dy, cout, x_raw, cin, 1)
            k.conv(cv[conv_pfx + "sc.dgrad"])
        self._wgrad(ws, pfx + "conv2.weight", k.KIND_CONV, dy, cout, a2, cout, 9)
        k.conv(cv[conv_pfx + "c2.dgrad"])
        self._bn_bwd(ws, sites[1], g_a2, h_raw, 0, None, 0, g_h, 0)
        self._wgrad(ws, pfx + "conv1.weight", k.KIND_CONV, g_h, cout, x_act, cin, 9)
        k.conv(cv[conv_pfx + "c1.dgrad"])
        self._bn_bwd(ws, sites[0], g_xact, x_raw, 0, g_sc if has_sc else dy, 0, g_x, 0)

    @_on_device
    def backward(self, dwave, async_allreduce=None):
        """dwave (B, L) or (B, 1, L): gradient of the loss w.r.t. the last forward's waveform.  Fills the flat gradient
        buffer ``self.G`` (every live parameter's gradient is overwritten, not accumulated).  ``async_allreduce`` is called
        with (lo, hi) element ranges of G as soon as they are final (bucket A after the decoder, bucket B at the end)."""
        k, ws = self.k, self._last
        B, L = ws.B, ws.L
        base = self.model.base
        hi, lo, window, tw = self._spectral_tables()
        dwave = dwave.detach().to(torch.float32).reshape(B, L).contiguous()
        if self._p2p_active(ws):
            ws.p2p.views[ws.epoch & 1].flat.zero_()
        else:
            ws.bsums_flat.zero_()
        # the weight-gradient and bias-sum launches ADD into the gradient buffers: one memset each per step instead of one per
        # launch (a launch behind a memset node cannot overlap its predecessor's tail)
        self.G[:self.live_end].zero_()
        self.Gp.zero_()
        k.istft_bwd(dwave, window, hi, lo, self.n_fft, self.hop, ws.T, ws.stft_ws, ws.dre, ws.dim)
        k.mask_bwd(ws.feat, ws.mag, ws.cos, ws.sin, ws.dre, ws.dim, ws.dfeat, self.n_fft)
        k.after_bwd(ws.dfeat, ws.d_raw[0], self._after_w, ws.g_y[0], self.g("after.w").view(3, 32), self.g("after.b"))
        cv = ws.conv
        for j in reversed(range(6)):
            cin, cout, (uh, uw) = DEC[j]
            lin, lo_ = 6 - j, 5 - j
            s0 = 14 + 3 * j
            self._block_bwd(ws, "dec%d.cb2." % j, (s0 + 1, s0 + 2), ws.g_y[lo_], ws.cat_raw[lo_], ws.cat_act[lo_], ws.hd_raw[lo_],
                            ws.a2d[lo_], ws.g_a2[lo_], ws.g_h[lo_], ws.g_catact[lo_], ws.g_sc_d[lo_], ws.g_cat[lo_], 2 * cout,
                            cout, True, "dec%d." % j)
            k.unshuffle(ws.g_cat[lo_], 0, cout, ws.dU[lin], uh, uw)
            self._wgrad(ws, "dec%d.up" % j, k.KIND_CONVT, ws.dU[lin], uh * uw * cout, ws.xin_act[lin], cin, 1)
            k.conv(cv["dec%d.up.dgrad" % j])
            self._bn_bwd(ws, s0, ws.g_xinact[lin], ws.d_raw[lin], 0, None, 0, ws.g_y[lin], 0)
        k.unpack_grads(self._unpack_a)
        if async_allreduce is not None:
            async_allreduce(0, self.bucket_a_end)
        for kk in reversed(range(7)):
            cin, cout, pool = ENC[kk]
            if kk < 6:
                k.pool_bwd(ws.g_xraw[kk + 1], ws.g_cat[kk], cout, ws.g_y[kk], pool[0], pool[1])
            self._block_bwd(ws, "enc%d." % kk, (2 * kk, 2 * kk + 1), ws.g_y[kk], ws.x_raw[kk], ws.x_act[kk], ws.h_raw[kk],
                            ws.a2[kk], ws.g_a2[kk], ws.g_h[kk], ws.g_xact[kk], ws.g_sc_e[kk], ws.g_xraw[kk], cin, cout,
                            cin != cout, "enc%d." % kk)
            if kk == _B1_LAST_LEVEL:
                k.unpack_grads(self._unpack_b1)
                if async_allreduce is not None:
                    async_allreduce(self.bucket_a_end, self.bucket_b1_end)
        k.unpack_grads(self._unpack_b2)
        k.pre_bwd(ws.g_xraw[0], ws.mag, self.bnp0, base.pre_conv.weight.data.view(32), self.g("pre.w").view(32),
                  self.g("pre.b"), self.g("bn0.weight"), self.g("bn0.bias"))
        k.film_bwd(ws.dbeta, ws.cond, self.G[self.film_w_off:self.film_w_off + self.J * self.K].view(self.J, self.K),
                   self.G[self.film_b_off:self.film_b_off + self.J])
        if async_allreduce is not None:
            async_allreduce(self.bucket_b1_end, self.live_end)

    # ------------------------------------------------------------------ fused step (loss + backward + all-reduce + AdamW)
    @_on_device
    def training_step(self, mixture, condition, target, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                      process_group=None):
        """One optimisation step (reference models/audiosep.py:52-145 + DDP): returns the loss as a 0-d tensor (this rank's
        mean |output - target|).  With torch.distributed initialised the gradients are summed over ranks by NCCL in two
        buckets — the first overlaps the encoder's backward — and divided by the world size inside the optimizer kernel."""
        import torch.distributed as dist
        k = self.k
        if process_group is not None:
            self.process_group = process_group
        wave = self.forward(mixture, condition)
        ws = self._last
        ws.loss_sum.zero_()
        k.l1_loss(ws.wave, target.detach().to(torch.float32).reshape(ws.B, ws.L).contiguous(), ws.dwave, ws.loss_sum)
        world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        works = []
        if world > 1 and self.device.type == "cuda":
            comm = self._collective_stream()

            def launch(lo_, hi_):
                ev = torch.cuda.Event()
                ev.record()
                comm.wait_event(ev)
                with torch.cuda.stream(comm):
                    dist.all_reduce(self.G[lo_:hi_], group=process_group)
            self.backward(ws.dwave, launch)
            torch.cuda.current_stream().wait_stream(comm)
        elif world > 1:
            self.backward(ws.dwave)
            dist.all_reduce(self.G[:self.live_end], group=process_group)
        else:
            self.backward(ws.dwave)
        del works
        self.optimizer_step(lr, betas, eps, weight_decay, grad_scale=1.0 / world)
        return ws.loss_sum[0] / float(ws.B * ws.L)

    @_on_device
    def optimizer_step(self, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
        """Fused AdamW(amsgrad=True) over the live slice of the flat buffers, then refresh the 16-bit weight layouts."""
        if self.opt_state is None:
            self.opt_state = [torch.zeros_like(self.P) for _ in range(3)]
        m, v, vmax = self.opt_state
        self.step_count += 1
        n = self.live_end
        self.k.adamw_amsgrad(self.P[:n], self.G[:n], m[:n], v[:n], vmax[:n], lr, betas[0], betas[1], eps, weight_decay,
                             self.step_count, grad_scale)
        self.refresh_weights()

    def grads(self):
        """{parameter: gradient view into the flat buffer} for every live parameter (dead ones are absent)."""
        out = {}
        for name, (off, p) in self.index.items():
            if not name.startswith("dead."):
                out[p] = self.G[off:off + p.numel()].view(p.shape)
        return out


class _TrainForward(torch.autograd.Function):
    """Autograd bridge: ``ResUNet30.forward`` in train mode returns a waveform connected to the parameters, so the
    reference's ``loss.backward(); optimizer.step()`` works unchanged (models/audiosep.py:100-111)."""

    @staticmethod
    def forward(ctx, engine, mixture, condition, *params):
        ctx.engine = engine
        return engine.forward(mixture, condition).clone()

    @staticmethod
    def backward(ctx, dwave):
        eng = ctx.engine
        eng.backward(dwave.contiguous())
        flat = eng.G.clone()
        grads = []
        for name, (off, p) in eng.index.items():
            grads.append(None if name.startswith("dead.") else flat[off:off + p.numel()].view(p.shape))
        return (None, None, None) + tuple(grads)


def train_forward(engine, mixture, condition):
    params = [p for _name, (_off, p) in engine.index.items()]
    return _TrainForward.apply(engine, mixture, condition, *params)
