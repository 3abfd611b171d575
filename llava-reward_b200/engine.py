"""RewardEngine: the scoring forward (reference `CustomRewardModel.custom_forward`, phi3v branch,
llava_reward/models/rw_model_general_preference.py:334-448) as a fixed sequence of C-ABI kernel launches.

Host side is plumbing only: buffer allocation (torch.empty), one tiny D2H read of per-sample counts at
the top of the call (the reference performs ~70 hidden syncs per forward), pointer passing.
Every FLOP and every byte moved on the device is done by libllavareward.so.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib as L
from . import ops
from .config import RewardConfig, num_image_tokens, packed_row_plan, qwen_window_plan
from .weights import PackedWeights


class RewardEngine:
    def __init__(self, cfg: RewardConfig, weights: PackedWeights, device="cuda", gemm_impl: int = L.GEMM_TCGEN05,
                 attn_impl: int = L.ATTN_TCGEN05, precision: str = "bf16"):
        L.load()
        self.cfg, self.w, self.device, self.gemm_impl = cfg, weights, torch.device(device), gemm_impl
        self.attn_impl = attn_impl
        # "fp32" = the verification path (csrc/f32_verify.cu): the same dataflow, layouts, plans and index kernels with
        # every floating-point kernel in plain fp32 and no bf16 rounding point; the weights must have been packed with
        # pack_weights(dtype=torch.float32). Debug configuration for the 1e-4 parity gate, never used for scoring.
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"precision {precision!r}: expected 'bf16' or 'fp32'")
        self.precision = precision
        self.dtype = torch.float32 if precision == "fp32" else torch.bfloat16
        self._bufs: Dict[Tuple[str, Tuple[int, ...], torch.dtype], torch.Tensor] = {}
        self._rope: Dict[Tuple[int, bool], Tuple[torch.Tensor, torch.Tensor]] = {}
        self._pinned: Dict[str, torch.Tensor] = {}
        self.launches = 0        # C-ABI calls issued by the last forward (= kernel launches, see _lib.launch_count)
        self.taps: Optional[dict] = None  # set to {} to capture intermediates (tests)
        self.profile: Optional[dict] = None  # {"gate_up": []} -> CUDA-event pairs around that GEMM (bench roofline)
        # Output-identical shortcut: after the attention of the LAST decoder layer only the
        # last-valid-token row of every sample is read by the head, so o_proj / post-norm / MLP of that layer run on
        # those B rows. Off while intermediates are being captured (taps) so tests can compare both forms.
        self.last_layer_rows = True
        # Output-identical shortcut: the decoder runs on the VALID rows only, packed back to
        # back (padded positions never reach the head; every kernel of the decoder is row-wise except the attention,
        # which already works on [start, start+len) of each sample and has a packed-sequence layout). Plain eval path
        # with a tcgen05 attention only; off while intermediates are captured.
        self.pack_rows = True

    # ------------------------------------------------------------------ helpers
    def buf(self, name: str, shape, dtype=None) -> torch.Tensor:
        dtype = self.dtype if dtype is None else dtype
        key = (name, tuple(shape), dtype)
        t = self._bufs.get(key)
        if t is None:
            # drop older shapes of the same logical buffer
            for k in [k for k in self._bufs if k[0] == name]:
                del self._bufs[k]
            t = torch.empty(*shape, dtype=dtype, device=self.device)
            self._bufs[key] = t
        return t

    def _h2d(self, name: str, host: torch.Tensor, dev: torch.Tensor) -> None:
        """dev.copy_(host) through a persistent pinned staging buffer (grown on demand) instead of a fresh
        `pin_memory()` allocation per call. Safe to reuse: every forward starts with a blocking D2H read of the token
        plan, so the previous forward's copy out of this buffer has completed before it is rewritten."""
        n = host.numel()
        st = self._pinned.get(name)
        if st is None or st.numel() < n or st.dtype != host.dtype:
            st = torch.empty(max(n, 1) * 2, dtype=host.dtype).pin_memory()
            self._pinned[name] = st
        st[:n].copy_(host.reshape(-1))
        dev.reshape(-1)[:n].copy_(st[:n], non_blocking=True)

    def rope_tables(self, n_pos: int, long: bool):
        """cos/sin [n_pos, head_dim/2] bf16, computed like Phi3SuScaledRotaryEmbedding.forward
        (modeling_phi3_v.py:446-476): fp32 angles, * scaling factor, cast to bf16."""
        key = (n_pos, long)
        if key not in self._rope:
            cfg = self.cfg
            fac = cfg.long_factor if long else cfg.short_factor
            ext = torch.tensor(fac, dtype=torch.float32, device=self.device)
            expo = torch.arange(0, cfg.head_dim, 2, dtype=torch.int64, device=self.device).float() / cfg.head_dim
            inv_freq = 1.0 / (ext * cfg.rope_theta ** expo)
            ang = torch.arange(n_pos, dtype=torch.int64, device=self.device).float()[:, None] * inv_freq[None, :]
            s = cfg.rope_scaling_factor
            self._rope = {key: ((ang.cos() * s).to(self.dtype).contiguous(),
                                (ang.sin() * s).to(self.dtype).contiguous())}
        return self._rope[key]

    def _gemm(self, A, W, C, M, N, K, epi=L.EPI_NONE, bias=None, R=None):
        ops.gemm(A, W, C, M, N, K, epi, bias, R, impl=self.gemm_impl)

    def _tap(self, name, t):
        if self.taps is not None:
            self.taps[name] = t.clone()

    def _clip_tower(self, pix: torch.Tensor, crop_idx: torch.Tensor, n_crops: int) -> torch.Tensor:
        """CLIP ViT-L/14-336 up to encoder layer `clip_layers` on the crops pix.view(-1,3,336,336)[crop_idx]
        -> tokens [n_crops*577, 1024] (row 0 of every crop = CLS)."""
        cfg, w = self.cfg, self.w
        D, DI, T = cfg.clip_hidden, cfg.clip_intermediate, cfg.clip_tokens
        Mv = n_crops * T
        a0 = self.buf("clip_a0", (n_crops * (T - 1), 640))
        ops.clip_im2col(pix, crop_idx, a0, n_crops)
        patch = self.buf("clip_patch", (n_crops * (T - 1), D))
        self._gemm(a0, w.clip["patch_w"], patch, n_crops * (T - 1), D, 640)
        x = self.buf("clip_x", (Mv, D))
        ops.clip_embed_ln(patch, w.clip["cls"], w.clip["pos"], w.clip["pre_w"], w.clip["pre_b"], x, n_crops, cfg.clip_eps)
        self._tap("clip_embed", x)
        hn = self.buf("clip_hn", (Mv, D))
        qkv = self.buf("clip_qkv", (Mv, 3 * D))
        ao = self.buf("clip_ao", (Mv, D))
        ff = self.buf("clip_ff", (Mv, DI))
        scale = cfg.clip_head_dim ** -0.5
        for li, lw in enumerate(w.clip_layers):
            ops.layernorm(x, lw["ln1_w"], lw["ln1_b"], hn, Mv, D, cfg.clip_eps)
            self._gemm(hn, lw["qkv_w"], qkv, Mv, 3 * D, D, L.EPI_BIAS, lw["qkv_b"])
            ops.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], ao, 3 * D, D, n_crops, T, None, None, cfg.clip_heads,
                          cfg.clip_head_dim, False, scale, self.attn_impl)
            self._gemm(ao, lw["out_w"], x, Mv, D, D, L.EPI_BIAS_RESIDUAL, lw["out_b"], x)
            ops.layernorm(x, lw["ln2_w"], lw["ln2_b"], hn, Mv, D, cfg.clip_eps)
            self._gemm(hn, lw["fc1_w"], ff, Mv, DI, D, L.EPI_BIAS_QUICKGELU, lw["fc1_b"])
            self._gemm(ff, lw["fc2_w"], x, Mv, D, DI, L.EPI_BIAS_RESIDUAL, lw["fc2_b"], x)
            if li == 0:
                self._tap("clip_layer0", x)
        self._tap("clip_out", x)
        return x

    def _decoder(self, hid, B: int, S: int, pos, seq_start, seq_len, cos_tab, sin_tab, eos_row=None, n_layers=None,
                 packed=None):
        """Pre-norm decoder layers in place on hid [B*S, H] (Phi3DecoderLayer, modeling_phi3_v.py:1130-1205; the
        Llama layers of the LLaVA-v1.6 branch have the same dataflow). LoRA ranks come from the packed weights:
        the K-extension of a fused projection is the stack of its branches' ranks (weights.py).
        Returns None (the output is hid) or, with `eos_row` and the last-layer shortcut, the [B, H] output rows of the
        last layer at those row indices (hid then holds the INPUT of the last layer). `n_layers` stops after that many
        layers (the reference's `layer_id` attribute: hidden_states[layer_id], rw_model_general_preference.py:349-352)."""
        cfg, w = self.cfg, self.w
        layers = w.layers if n_layers is None else w.layers[:n_layers]
        H, I = cfg.hidden_size, cfg.intermediate_size
        # packed = (seq_base [B] int32, rows, longest sequence): hid holds the valid rows of sample b at
        # [seq_base[b], +seq_len[b]); buffers keep the B*S capacity so ragged batches do not re-allocate
        cap = B * S
        M = packed[1] if packed is not None else cap
        lw0 = w.layers[0] if w.layers else {}
        rq, ro, rg, rd = (lw0[k].shape[0] if k in lw0 else 0 for k in ("qkv_a", "o_a", "gu_a", "dn_a"))
        xn = self.buf("dec_xn", (cap, H + max(rq, rg)))[:M]
        nh, hd = cfg.num_heads, cfg.head_dim
        nkv = getattr(cfg, "num_kv_heads", nh)      # grouped-query attention (Qwen2 decoder): k/v are nkv*hd wide
        kvw = nkv * hd
        QW = H + 2 * kvw
        dqkv = self.buf("dec_qkv", (cap, QW))[:M]
        dao = self.buf("dec_ao", (cap, H + ro))[:M]
        gg = self.buf("dec_g", (cap, I + rd))[:M]
        att_scale = 1.0 / math.sqrt(cfg.head_dim)
        for li, lw in enumerate(layers):
            ops.rmsnorm(hid, lw["in_ln"], xn, M, H, cfg.rms_eps)
            if rq:
                self._gemm(xn, lw["qkv_a"], xn[:, H:], M, rq, H)
            # qkv projection with RoPE in the epilogue (q/k rows are head-interleaved, see weights.py)
            if "qkv_b" in lw:   # q/k/v biases (Qwen2): bias + RoPE epilogue, per-token tables when pos is None
                ops.gemm_rope_ex(xn, lw["qkv_w"], dqkv, M, QW, H + rq, lw["qkv_b"], pos, cos_tab, sin_tab, H + kvw, hd,
                                 L.EPI_BIAS_ROPE)
            else:
                ops.gemm_rope(xn, lw["qkv_w"], dqkv, M, QW, H + rq, pos, cos_tab, sin_tab, H + kvw, hd,
                              L.GEMM_SIMT if self.gemm_impl == L.GEMM_SIMT else L.GEMM_TCGEN05)
            if packed is not None:
                ops.attention_ex(dqkv, dqkv[:, H:], dqkv[:, H + kvw:], dao, QW, H + ro, M, B, packed[2], packed[0], None,
                                 seq_len, nh, nkv, hd, True, att_scale, self.attn_impl)
            elif nkv != nh:
                ops.attention_ex(dqkv, dqkv[:, H:], dqkv[:, H + kvw:], dao, QW, H + ro, M, B, S, None, seq_start,
                                 seq_len, nh, nkv, hd, True, att_scale, self.attn_impl)
            else:
                ops.attention(dqkv, dqkv[:, H:], dqkv[:, 2 * H:], dao, 3 * H, H + ro, B, S, seq_start, seq_len, nh, hd,
                              True, att_scale, self.attn_impl)
            if eos_row is not None and self.last_layer_rows and self.taps is None and li == len(layers) - 1:
                # last layer: everything after the attention on the B last-valid-token rows only
                dao_e = self.buf("dec_ao_e", (B, H + ro))
                hid_e = self.buf("dec_hid_e", (B, H))
                ops.gather_rows(dao, eos_row, dao_e, B, H)
                ops.gather_rows(hid, eos_row, hid_e, B, H)
                if ro:
                    self._gemm(dao_e, lw["o_a"], dao_e[:, H:], B, ro, H)
                self._gemm(dao_e, lw["o_w"], hid_e, B, H, H + ro, L.EPI_RESIDUAL, None, hid_e)
                xn_e = self.buf("dec_xn_e", (B, H + rg))
                ops.rmsnorm(hid_e, lw["post_ln"], xn_e, B, H, cfg.rms_eps)
                if rg:
                    self._gemm(xn_e, lw["gu_a"], xn_e[:, H:], B, rg, H)
                gg_e = self.buf("dec_g_e", (B, I + rd))
                self._gemm(xn_e, lw["gu_w"], gg_e, B, 2 * I, H + rg, L.EPI_SWIGLU)
                if rd:
                    self._gemm(gg_e, lw["dn_a"], gg_e[:, I:], B, rd, I)
                self._gemm(gg_e, lw["dn_w"], hid_e, B, H, I + rd, L.EPI_RESIDUAL, None, hid_e)
                return hid_e
            if ro:
                self._gemm(dao, lw["o_a"], dao[:, H:], M, ro, H)
            self._gemm(dao, lw["o_w"], hid, M, H, H + ro, L.EPI_RESIDUAL, None, hid)
            ops.rmsnorm(hid, lw["post_ln"], xn, M, H, cfg.rms_eps)
            if rg:
                self._gemm(xn, lw["gu_a"], xn[:, H:], M, rg, H)
            if self.profile is not None:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            self._gemm(xn, lw["gu_w"], gg, M, 2 * I, H + rg, L.EPI_SWIGLU)
            if self.profile is not None:
                ev[1].record()
                self.profile["gate_up"].append(ev + (M,))   # rows of this launch (packed valid rows or B*S)
            if rd:
                self._gemm(gg, lw["dn_a"], gg[:, I:], M, rd, I)
            self._gemm(gg, lw["dn_w"], hid, M, H, I + rd, L.EPI_RESIDUAL, None, hid)
            if self.taps is not None:
                self._tap(f"hidden_{li}", hid)
        return None

    def _final_rows(self, hid, hid_e, eos_row, B: int, final_norm: bool = True) -> torch.Tensor:
        """final RMSNorm of the last-valid-token rows -> x_eos [B, H] (`final_norm` False: the rows themselves, for
        the un-normed hidden_states[layer_id] of the reference's `layer_id` attribute)"""
        cfg, w = self.cfg, self.w
        xe = self.buf("x_eos", (max(B, 1), cfg.hidden_size))
        if not final_norm:
            if hid_e is not None:
                return hid_e
            ops.gather_rows(hid, eos_row, xe, B, cfg.hidden_size)
            return xe
        if hid_e is not None:
            ops.rmsnorm(hid_e, w.head["norm"], xe, B, cfg.hidden_size, cfg.rms_eps)
        else:
            ops.rmsnorm(hid, w.head["norm"], xe, B, cfg.hidden_size, cfg.rms_eps, row_index=eos_row)
        self._tap("last_hidden_eos", xe)
        return xe

    def model_outputs(self, taps: dict, B: int, S: int):
        """The `outputs` object of custom_forward(return_output=True) from a forward run with `self.taps` set:
        last_hidden_state = norm(h_L) and hidden_states = (inputs_embeds, h_1 .. h_{L-1}, norm(h_L), vision_embeds),
        vision_embeds zero-padded to the batch maximum (modeling_phi3_v.py:250-252, 1463-1505). Rows at padded
        positions follow the reference's flash-attention path (attention output 0 there), not its eager path."""
        from transformers.modeling_outputs import BaseModelOutputWithPast

        cfg, w = self.cfg, self.w
        H, n = cfg.hidden_size, cfg.num_layers
        last = torch.empty(B * S, H, dtype=self.dtype, device=self.device)
        src = taps[f"hidden_{n - 1}"] if n > 0 else taps["inputs_embeds"]
        ops.rmsnorm(src, w.head["norm"], last, B * S, H, cfg.rms_eps)
        plan_h, max_nv = self._last_plan
        idx_h = np.full((B, max_nv), -1, dtype=np.int32)
        for b in range(B):
            nv = int(plan_h[b, L.PLAN_NV])
            idx_h[b, :nv] = int(plan_h[b, L.PLAN_ROW_BASE]) + np.arange(nv, dtype=np.int32)
        idx = torch.from_numpy(idx_h.reshape(-1)).to(self.device)
        vis = torch.empty(B * max_nv, H, dtype=self.dtype, device=self.device)
        ops.gather_rows(taps["img_proj"], idx, vis, B * max_nv, H)
        hs = [taps["inputs_embeds"]] + [taps[f"hidden_{i}"] for i in range(n - 1)] + [last]
        hs = tuple(t.view(B, S, H) for t in hs) + (vis.view(B, max_nv, H),)
        return BaseModelOutputWithPast(last_hidden_state=hs[-2], past_key_values=None, hidden_states=hs, attentions=None)

    def lm_outputs(self, taps: dict, B: int, S: int):
        """The `outputs` object of custom_forward(return_output=True) for the llava / qwen branches (the reference
        returns transformers' CausalLMOutputWithPast of `self.forward(**inputs_batch, output_hidden_states=True)`,
        rw_model_general_preference.py:357, 374): hidden_states = (inputs_embeds, h_1 .. h_{L-1}, norm(h_L)) from a
        forward run with `self.taps` set. `logits` is None: the vocabulary-wide lm_head GEMM the reference computes and
        custom_forward never reads is not executed here."""
        from transformers.modeling_outputs import CausalLMOutputWithPast

        cfg, w = self.cfg, self.w
        H, n = cfg.hidden_size, cfg.num_layers
        last = torch.empty(B * S, H, dtype=torch.bfloat16, device=self.device)
        ops.rmsnorm(taps[f"hidden_{n - 1}"] if n > 0 else taps["inputs_embeds"], w.head["norm"], last, B * S, H, cfg.rms_eps)
        hs = [taps["inputs_embeds"]] + [taps[f"hidden_{i}"] for i in range(n - 1)] + [last]
        return CausalLMOutputWithPast(loss=None, logits=None, past_key_values=None,
                                      hidden_states=tuple(t.view(B, S, H) for t in hs), attentions=None)

    def _pack_valid_rows(self, hid, meta_h, B: int, S: int, pos_from_zero: bool):
        """Gather the valid rows of hid [B*S, H] back to back
        -> (hid_p [rows, H], pos_p, seq_base, eos_p, rows, longest, row index of every packed row).
        Positions of the valid tokens are 0..len-1 for position_ids = cumsum(mask) - 1 (phi3v) and the slot index for
        position_ids = arange(S) (llava); the index / position vectors are built on the host from the token-plan
        record that is already there and go up in one copy."""
        H = self.cfg.hidden_size
        idx, posv, base, last = packed_row_plan(meta_h[:B], meta_h[B:2 * B], S, pos_from_zero)
        rows = int(idx.shape[0])
        host = torch.from_numpy(np.concatenate([idx, posv, base, last]))
        dev = self.buf("pack_plan", (2 * B * S + 2 * B,), torch.int32)[: host.numel()]
        self._h2d("pack_plan", host, dev)
        idx_d, pos_d = dev[:rows], dev[rows:2 * rows]
        base_d, eos_d = dev[2 * rows:2 * rows + B], dev[2 * rows + B:]
        hid_p = self.buf("hidden_packed", (B * S, H))[:rows]
        ops.gather_rows(hid, idx_d, hid_p, rows, H)
        return hid_p, pos_d, base_d, eos_d, rows, int(meta_h[B:2 * B].max()), idx_d

    def _can_pack(self, meta_h, B: int, S: int, last_position: bool, mean_pool: bool) -> bool:
        # a sample with an all-zero attention mask has no rows to pack (packed_row_plan would point its head row at the
        # previous sample): the slot layout then reads row S-1 of that sample, as the reference does
        return (self.pack_rows and self.taps is None and not last_position and not mean_pool and
                self.attn_impl != L.ATTN_MMA_SYNC and int(meta_h[B:2 * B].sum()) < B * S and
                int(meta_h[B:2 * B].min()) > 0)

    def _resolve_layer_id(self, layer_id):
        """-> (decoder layers to run, apply the final norm). The reference takes `last_hidden_state` for layer_id 32 and
        otherwise indexes (inputs_embeds, h_1 .. h_{L-1}, norm(h_L), vision_embeds) (rw_model_general_preference.py
        :349-352, modeling_phi3_v.py:1463-1505)."""
        n = self.cfg.num_layers
        if layer_id is None or layer_id == 32 or layer_id == n:
            return n, True
        if 0 <= layer_id < n:
            return int(layer_id), False
        raise ValueError(f"layer_id {layer_id}: expected 32 or 0..{n}")

    def _head_rows(self, eos_row, B: int, S: int, last_position: bool):
        """row index per sample the head reads: the last valid token (eval) or position S-1 (the reference's
        `self.training` gather, values[:, -1], rw_model_general_preference.py:413-418, 432-436)"""
        if not last_position:
            return eos_row
        key = ("last_rows", B, S)
        if getattr(self, "_last_rows_key", None) != key:
            self._last_rows = (torch.arange(B, dtype=torch.int32, device=self.device) * S + (S - 1)).contiguous()
            self._last_rows_key = key
        return self._last_rows

    def _mean_head(self, hid, mask, B: int, S: int, final_norm: bool, img=None, plan_h=None, max_nv: int = 0,
                   cross_attention=None, masked_pad: bool = False, ca_eps=None):
        """`mean_hidden_state` pooling (rw_model_general_preference.py:376-386 for ALL rows, then :398-406 and the value
        head): final norm of every row, the S x N_v cross attention of every sample as GEMMs on the zero-padded
        vision rows, residual + ca_layernorm, masked mean, value head.
        masked_pad = the qwen arm (:387-397): the padded vision rows are masked with -1e4 before the softmax (exp
        underflows to exactly 0, i.e. the softmax runs over the sample's own rows) instead of taking part in it with a
        score of 0 as in the phi3v arm; a sample without any vision row gets attn_o = 0 (uniform weights over zero
        rows of V). ca_eps: epsilon of ca_layernorm (the qwen loader builds it with 1e-6)."""
        cfg, w = self.cfg, self.w
        M, H = B * S, cfg.hidden_size
        if final_norm:
            xa = self.buf("x_all", (M, H))
            ops.rmsnorm(hid, w.head["norm"], xa, M, H, cfg.rms_eps)
        else:
            xa = hid
        self._tap("last_hidden_all", xa)
        ca_eps = cfg.rms_eps if ca_eps is None else ca_eps
        use_ca = cfg.add_cross_attention if cross_attention is None else cross_attention
        if use_ca and max_nv == 0:
            # no vision row in the whole batch (qwen arm): vision_pad is [B, 0, H], attn_o = 0 -> ca_layernorm(x + 0)
            y2 = self.buf("ca_y2", (M, H))
            ops.rmsnorm(xa, w.head["ca_ln"], y2, M, H, ca_eps)
            xa = y2
        elif use_ca:
            NVP = (max_nv + 255) // 256 * 256
            idx_h = np.full((B, NVP), -1, dtype=np.int32)  # -1 = the reference's zero-padded vision rows
            for b in range(B):
                nv = int(plan_h[b, L.PLAN_NV])
                idx_h[b, :nv] = int(plan_h[b, L.PLAN_ROW_BASE]) + np.arange(nv, dtype=np.int32)
            idx = self.buf("ca_pad_idx", (B * NVP,), torch.int32)
            idx.copy_(torch.from_numpy(idx_h.reshape(-1)))
            vis = self.buf("ca_vis_pad", (B * NVP, H))
            ops.gather_rows(img, idx, vis, B * NVP, H)
            wk, wv = w.head["wkv"][:H], w.head["wkv"][H:]
            kp = self.buf("ca_k_pad", (B * NVP, H))
            self._gemm(vis, wk, kp, B * NVP, H, H)
            q = self.buf("ca_q_all", (M, H))
            self._gemm(xa, w.head["wq"], q, M, H, H)
            sc = self.buf("ca_scores_all", (M, NVP))
            vt = self.buf("ca_vt", (B * H, NVP))      # V^T per sample: the K-major operand of the P.V GEMM
            y = self.buf("ca_y", (M, H))
            for b in range(B):
                self._gemm(q[b * S:(b + 1) * S], kp[b * NVP:(b + 1) * NVP], sc[b * S:(b + 1) * S], S, NVP, H)
                self._gemm(wv, vis[b * NVP:(b + 1) * NVP], vt[b * H:(b + 1) * H], H, NVP, H)
            if masked_pad:
                for b in range(B):
                    nv = int(plan_h[b, L.PLAN_NV])
                    if nv > 0:
                        ops.softmax_rows(sc[b * S:(b + 1) * S], S, nv, NVP, 1.0 / math.sqrt(H))
                    else:
                        sc[b * S:(b + 1) * S].zero_()
            else:
                ops.softmax_rows(sc, M, max_nv, NVP, 1.0 / math.sqrt(H))
            for b in range(B):
                self._gemm(sc[b * S:(b + 1) * S], vt[b * H:(b + 1) * H], y[b * S:(b + 1) * S], S, H, NVP,
                           L.EPI_RESIDUAL, None, xa[b * S:(b + 1) * S])
            y2 = self.buf("ca_y2", (M, H))
            ops.rmsnorm(y, w.head["ca_ln"], y2, M, H, ca_eps)
            xa = y2
            self._tap("skipca_all", xa)
        pooled = self.buf("pooled", (B, H))
        ops.masked_mean_rows(xa, mask, pooled, B, S, H)
        reward = torch.empty(B, cfg.vhd, dtype=self.dtype, device=self.device)
        ops.skipca_head(None, None, None, pooled, None, w.head["vh"], reward, B, H, 0, cfg.vhd, cfg.rms_eps)
        return reward

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, attention_mask: torch.Tensor, pixel_values: torch.Tensor,
                image_sizes, layer_id=None, last_position: bool = False, mean_pool: bool = False,
                vision_layer_id: int = -1) -> torch.Tensor:
        """layer_id / last_position / mean_pool / vision_layer_id = the reference's `layer_id`, `training`,
        `mean_hidden_state` and `vision_layer_id` attributes (rw_model_general_preference.py:327-333, 349-353); the
        defaults are the eval-mode scoring path."""
        cfg, w, dev = self.cfg, self.w, self.device
        bf = self.dtype
        if self.precision == "fp32" and mean_pool:
            raise NotImplementedError("mean_hidden_state in the fp32 verification path")
        n_run, final_norm = self._resolve_layer_id(layer_id)
        vis_from_layer = None
        if cfg.add_cross_attention and vision_layer_id not in (-1, cfg.num_layers + 1):
            # SkipCA keys/values = hidden_states[vision_layer_id][:, :N_v_max] instead of the projected image tokens
            # (:353). Analysis mode: every layer output is captured like for return_output.
            if mean_pool or n_run != cfg.num_layers:
                raise NotImplementedError("vision_layer_id together with mean_hidden_state / an early layer_id")
            vis_from_layer = range(cfg.num_layers + 2)[vision_layer_id]   # python indexing of the hidden_states tuple
            own_taps = self.taps is None
            if own_taps:
                self.taps = {}
        launches0 = L.launch_count()
        B, S = input_ids.shape
        H = cfg.hidden_size
        ids = input_ids.to(dev, torch.int64).contiguous()
        mask = attention_mask.to(dev, torch.int64).contiguous()
        pix = pixel_values.to(dev, torch.float32).contiguous()
        if pix.dim() != 5 or tuple(pix.shape[2:]) != (3, cfg.image_size, cfg.image_size):
            raise AssertionError("pixel_values must be [B, crops, 3, 336, 336]")  # modeling_phi3_v.py:236
        n_slots = pix.shape[1]

        # 1. token plan on device, then ONE small D2H read
        M = B * S
        pos = self.buf("pos", (M,), torch.int32)
        img_ord = self.buf("img_ord", (M,), torch.int32)
        meta = self.buf("meta", (4 * B + 1,), torch.int32)  # seq_start | seq_len | eos_row | n_img | flags
        seq_start, seq_len, eos_row, n_img, flags = meta[:B], meta[B:2 * B], meta[2 * B:3 * B], meta[3 * B:4 * B], meta[4 * B:]
        flags.zero_()
        ops.token_plan(ids, mask, B, S, pos, img_ord, seq_start, seq_len, eos_row, n_img, flags)
        meta_h = meta.cpu().numpy()
        sizes_h = image_sizes.cpu().numpy() if torch.is_tensor(image_sizes) else np.asarray(image_sizes)
        sizes_h = sizes_h.reshape(B, 2).astype(np.int64)
        if meta_h[4 * B] & 1:
            raise ValueError("attention_mask rows must be one contiguous run of ones (left/right padding)")
        plan_h = np.zeros((B, L.PLAN_STRIDE), dtype=np.int32)
        crop_src, crop_base, row_base = [], 0, 0
        for b in range(B):
            hc, wc = int(sizes_h[b, 0]) // 336, int(sizes_h[b, 1]) // 336
            ncrop = hc * wc + 1
            nv = num_image_tokens(int(sizes_h[b, 0]), int(sizes_h[b, 1]))
            if ncrop > n_slots:
                raise ValueError(f"image_sizes[{b}] needs {ncrop} crops but pixel_values has {n_slots} slots")
            if int(meta_h[3 * B + b]) != nv:
                # the reference fails in index_put with a shape mismatch (modeling_phi3_v.py:247-249)
                raise ValueError(f"sample {b}: {int(meta_h[3 * B + b])} image placeholder tokens but image_sizes "
                                 f"implies {nv}")
            plan_h[b, :5] = (hc, wc, crop_base, row_base, nv)
            crop_src.extend(range(b * n_slots, b * n_slots + ncrop))
            crop_base += ncrop
            row_base += nv
        n_crops, sum_nv = crop_base, row_base
        max_nv = int(plan_h[:, L.PLAN_NV].max())
        self._last_plan = (plan_h, max_nv)
        max_len = int(meta_h[B:2 * B].max())
        host = torch.from_numpy(np.concatenate([plan_h.reshape(-1), np.asarray(crop_src, dtype=np.int32)]))
        dev_plan = self.buf("plan", (host.numel(),), torch.int32)
        self._h2d("plan", host, dev_plan)
        plan = dev_plan[: B * L.PLAN_STRIDE]
        crop_idx = dev_plan[B * L.PLAN_STRIDE:]

        # 2. CLIP tower on the real crops only
        x = self._clip_tower(pix, crop_idx, n_crops)
        D = cfg.clip_hidden

        # 3. HD transform gather + projector
        rows = self.buf("hd_rows", (sum_nv, 4 * D))
        ops.hd_gather(x, plan, w.proj["sub_gn"], w.proj["glb_gn"], rows, B, max_nv)
        p1 = self.buf("proj1", (sum_nv, H))
        self._gemm(rows, w.proj["p0_w"], p1, sum_nv, H, 4 * D, L.EPI_BIAS_GELU, w.proj["p0_b"])
        img = self.buf("img_proj", (sum_nv, H))
        self._gemm(p1, w.proj["p2_w"], img, sum_nv, H, H, L.EPI_BIAS, w.proj["p2_b"])
        self._tap("img_proj", img)

        # 4. embeddings
        hid = self.buf("hidden", (M, H))
        ops.embed_scatter(ids, img_ord, plan, w.embed, img, hid, B, S, H, cfg.vocab_size)
        self._tap("inputs_embeds", hid)

        # 5. decoder
        cos_tab, sin_tab = self.rope_tables(max(S, 2), max_len > cfg.original_max_position_embeddings)
        head_row = self._head_rows(eos_row, B, S, last_position)
        if self._can_pack(meta_h, B, S, last_position, mean_pool):
            hid, pos_p, base_p, head_row, rows_p, longest, _ = self._pack_valid_rows(hid, meta_h, B, S, True)
            hid_e = self._decoder(hid, B, S, pos_p, None, seq_len, cos_tab, sin_tab, head_row, n_layers=n_run,
                                  packed=(base_p, rows_p, longest))
        else:
            hid_e = self._decoder(hid, B, S, pos, seq_start, seq_len, cos_tab, sin_tab,
                                  None if mean_pool else head_row, n_layers=n_run)
        if mean_pool:
            reward = self._mean_head(hid, mask, B, S, final_norm, img, plan_h, max_nv)
            self.launches = L.launch_count() - launches0
            return reward

        # 6. reward head on the last valid token of each sample
        xe = self._final_rows(hid, hid_e, head_row, B, final_norm)
        vhd = cfg.vhd
        reward = torch.empty(B, vhd, dtype=bf, device=dev)
        if cfg.add_cross_attention and vis_from_layer is not None:
            # rows [0, N_v_max) of every sample of the chosen hidden state, as the reference slices them (:353); all of
            # them are "real" rows for the softmax (no zero padding in this form)
            n = cfg.num_layers
            if vis_from_layer == 0:
                src = self.taps["inputs_embeds"]
            elif vis_from_layer < n:
                src = self.taps[f"hidden_{vis_from_layer - 1}"]
            else:
                src = torch.empty(M, H, dtype=bf, device=dev)
                ops.rmsnorm(self.taps[f"hidden_{n - 1}"], w.head["norm"], src, M, H, cfg.rms_eps)
            if own_taps:
                self.taps = None
            idx = (torch.arange(B, device=dev, dtype=torch.int32)[:, None] * S +
                   torch.arange(max_nv, device=dev, dtype=torch.int32)[None, :]).reshape(-1).contiguous()
            vis = self.buf("ca_vis_layer", (B * max_nv, H))
            ops.gather_rows(src, idx, vis, B * max_nv, H)
            plan2_h = np.zeros((B, L.PLAN_STRIDE), dtype=np.int32)
            plan2_h[:, L.PLAN_ROW_BASE] = np.arange(B, dtype=np.int32) * max_nv
            plan2_h[:, L.PLAN_NV] = max_nv
            plan2 = torch.from_numpy(plan2_h.reshape(-1)).to(dev)
            q = self.buf("ca_q", (B, H))
            self._gemm(xe, w.head["wq"], q, B, H, H)
            kv = self.buf("ca_kv_layer", (B * max_nv, 2 * H))
            self._gemm(vis, w.head["wkv"], kv, B * max_nv, 2 * H, H)
            scores = self.buf("ca_scores", (B, max_nv), torch.float32)
            ops.skipca_scores(q, kv, plan2, scores, B, H, max_nv)
            ops.skipca_head(scores, kv, plan2, xe, w.head["ca_ln"], w.head["vh"], reward, B, H, max_nv, vhd, cfg.rms_eps)
        elif cfg.add_cross_attention:
            q = self.buf("ca_q", (B, H))
            self._gemm(xe, w.head["wq"], q, B, H, H)
            kv = self.buf("ca_kv", (sum_nv, 2 * H))
            self._gemm(img, w.head["wkv"], kv, sum_nv, 2 * H, H)
            scores = self.buf("ca_scores", (B, max_nv), torch.float32)
            ops.skipca_scores(q, kv, plan, scores, B, H, max_nv)
            ops.skipca_head(scores, kv, plan, xe, w.head["ca_ln"], w.head["vh"], reward, B, H, max_nv, vhd, cfg.rms_eps)
        else:
            ops.skipca_head(None, None, None, xe, None, w.head["vh"], reward, B, H, 0, vhd, cfg.rms_eps)
        self.launches = L.launch_count() - launches0
        return reward

    @torch.no_grad()
    def preference(self, chosen: torch.Tensor, reject: torch.Tensor) -> torch.Tensor:
        """fp32 device tensor of P(chosen > reject) (preference_compute without the .cpu().numpy())."""
        cfg = self.cfg
        n = chosen.shape[0]
        prob = torch.empty(n, dtype=torch.float32, device=self.device)
        ops.preference(chosen.to(self.dtype).contiguous(), reject.to(self.dtype).contiguous(), prob, n,
                       chosen.shape[1], cfg.is_general_preference, cfg.general_preference_tau)
        return prob


class LlavaNextRewardEngine(RewardEngine):
    """The reference's llava branch (rw_model_general_preference.py:372-375 -> LlavaNextForConditionalGeneration.forward
    with output_hidden_states, value head on the last valid token :407-448) as C-ABI launches:
    CLIP (all real patches) -> projector on every CLIP token -> [embedding gather + anyres unpad/newline pack] in one
    kernel -> Llama decoder (same kernels as the Phi-3 loop: fused q/k/v + RoPE epilogue, tcgen05 attention hd 128,
    SwiGLU epilogue, LoRA K-extension) -> final RMSNorm on the last valid row -> value head.
    The reference also runs the 32064-wide lm_head over every token and discards it; that GEMM is not executed."""

    def rope_tables(self, n_pos: int, long: bool = False):
        """cos/sin [n_pos, 64] bf16 of LlamaRotaryEmbedding (default rope; transformers modeling_llama.py:124-135)."""
        key = (n_pos, False)
        if key not in self._rope:
            cfg = self.cfg
            expo = torch.arange(0, cfg.head_dim, 2, dtype=torch.int64, device=self.device).float() / cfg.head_dim
            inv_freq = 1.0 / (cfg.rope_theta ** expo)
            ang = torch.arange(n_pos, dtype=torch.int64, device=self.device).float()[:, None] * inv_freq[None, :]
            self._rope = {key: (ang.cos().to(torch.bfloat16).contiguous(), ang.sin().to(torch.bfloat16).contiguous())}
        return self._rope[key]

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, attention_mask: torch.Tensor, pixel_values: torch.Tensor,
                image_sizes, last_position: bool = False, mean_pool: bool = False) -> torch.Tensor:
        from .config import anyres_geometry

        cfg, w, dev = self.cfg, self.w, self.device
        launches0 = L.launch_count()
        B, S = input_ids.shape
        H = cfg.hidden_size
        ids = input_ids.to(dev, torch.int64).contiguous()
        mask = attention_mask.to(dev, torch.int64).contiguous()
        pix = pixel_values.to(dev, torch.float32).contiguous()
        if pix.dim() != 5 or tuple(pix.shape[2:]) != (3, cfg.image_size, cfg.image_size):
            # the 4-d "list of patches" form of get_image_features (modeling_llava_next.py:390-392) is not produced
            # by the reference's collate_fn (reward_dataset.py:334-346)
            raise ValueError(f"pixel_values of shape {tuple(pix.shape)}, expect [B, patches, 3, 336, 336]")
        n_slots = pix.shape[1]

        # 1. token plan (image positions = image_token_id, positions = arange) + ONE small D2H read
        M = B * S
        pos = self.buf("pos", (M,), torch.int32)
        img_ord = self.buf("img_ord", (M,), torch.int32)
        meta = self.buf("meta", (4 * B + 1,), torch.int32)
        seq_start, seq_len, eos_row, n_img, flags = meta[:B], meta[B:2 * B], meta[2 * B:3 * B], meta[3 * B:4 * B], meta[4 * B:]
        flags.zero_()
        ops.token_plan_ex(ids, mask, B, S, cfg.image_token_id, L.POS_ARANGE, pos, img_ord, seq_start, seq_len, eos_row,
                          n_img, flags)
        meta_h = meta.cpu().numpy()
        sizes_h = image_sizes.cpu().numpy() if torch.is_tensor(image_sizes) else np.asarray(image_sizes)
        sizes_h = sizes_h.reshape(B, 2).astype(np.int64)
        if meta_h[4 * B] & 1:
            raise ValueError("attention_mask rows must be one contiguous run of ones (left/right padding)")
        plan_h = np.zeros((B, L.PLAN_STRIDE), dtype=np.int32)
        patch_src, patch_base, row_base = [], 0, 0
        for b in range(B):
            g = anyres_geometry((int(sizes_h[b, 0]), int(sizes_h[b, 1])), cfg.image_grid_pinpoints, cfg.image_size,
                                cfg.patch)
            if g["n_patches"] > n_slots:
                raise ValueError(f"image_sizes[{b}] needs {g['n_patches']} patches but pixel_values has {n_slots}")
            if int(meta_h[3 * B + b]) != g["n_tokens"]:
                # transformers checks the batch total only (modeling_llava_next.py:437-441) and would shift features
                # across samples on a per-sample mismatch; that is never a valid input, so it is an error here
                raise ValueError(f"Image features and image tokens do not match, sample {b}: tokens: "
                                 f"{int(meta_h[3 * B + b])}, features: {g['n_tokens']}")
            plan_h[b, :7] = (g["grid_h"], g["grid_w"], patch_base, row_base, g["n_tokens"], g["top"], g["left"])
            patch_src.extend(range(b * n_slots, b * n_slots + g["n_patches"]))
            patch_base += g["n_patches"]
            row_base += g["n_tokens"]
        n_patches = patch_base
        host = torch.from_numpy(np.concatenate([plan_h.reshape(-1), np.asarray(patch_src, dtype=np.int32)]))
        dev_plan = self.buf("plan", (host.numel(),), torch.int32)
        self._h2d("plan", host, dev_plan)
        plan = dev_plan[: B * L.PLAN_STRIDE]
        patch_idx = dev_plan[B * L.PLAN_STRIDE:]

        # 2. CLIP tower, 3. projector on all 577 tokens of every patch (the CLS rows are computed and never read:
        #    0.17 % extra rows instead of a compaction pass)
        x = self._clip_tower(pix, patch_idx, n_patches)
        Mv = n_patches * cfg.clip_tokens
        p1 = self.buf("proj1", (Mv, H))
        self._gemm(x, w.proj["p0_w"], p1, Mv, H, cfg.clip_hidden, L.EPI_BIAS_GELU, w.proj["p0_b"])
        feat = self.buf("img_proj", (Mv, H))
        self._gemm(p1, w.proj["p2_w"], feat, Mv, H, H, L.EPI_BIAS, w.proj["p2_b"])
        self._tap("projector_out", feat)

        # 4. embeddings + anyres pack (unpad, image_newline) in one pass
        hid = self.buf("hidden", (M, H))
        ops.anyres_embed_scatter(ids, img_ord, plan, w.embed, feat, w.proj["newline"], hid, B, S, H, cfg.vocab_size)
        self._tap("inputs_embeds", hid)

        # 5. decoder
        cos_tab, sin_tab = self.rope_tables(max(S, 2))
        head_row = self._head_rows(eos_row, B, S, last_position)
        if self._can_pack(meta_h, B, S, last_position, mean_pool):
            # position_ids = arange(S) in this branch: a valid token keeps its slot index as position
            hid, pos_p, base_p, head_row, rows_p, longest, _ = self._pack_valid_rows(hid, meta_h, B, S, False)
            hid_e = self._decoder(hid, B, S, pos_p, None, seq_len, cos_tab, sin_tab, head_row,
                                  packed=(base_p, rows_p, longest))
        else:
            hid_e = self._decoder(hid, B, S, pos, seq_start, seq_len, cos_tab, sin_tab,
                                  None if mean_pool else head_row)
        if mean_pool:   # the reference has no SkipCA arm for this backbone: final norm, masked mean, value head
            reward = self._mean_head(hid, mask, B, S, True, cross_attention=False)
            self.launches = L.launch_count() - launches0
            return reward

        # 6. final norm on the last valid row + value head
        xe = self._final_rows(hid, hid_e, head_row, B)
        reward = torch.empty(B, cfg.vhd, dtype=torch.bfloat16, device=dev)
        ops.skipca_head(None, None, None, xe, None, w.head["vh"], reward, B, H, 0, cfg.vhd, cfg.rms_eps)
        self.launches = L.launch_count() - launches0
        return reward


class QwenVLRewardEngine(RewardEngine):
    """The reference's qwen branch (rw_model_general_preference.py:354-371 forward, :387-397 SkipCA, :407-448 head) as
    C-ABI launches: [patch rows -> bf16, window order] -> patch-embed GEMM -> 32 vision blocks (RMSNorm, qkv GEMM with
    bias + fp32 2D rotary in the epilogue, packed varlen tcgen05 attention over windows / whole images, proj GEMM,
    RMSNorm, gate|up GEMM with bias + SwiGLU epilogue, down GEMM) -> merger (RMSNorm that also undoes the window
    permutation, 2 GEMMs) -> embedding gather + image scatter -> Qwen2 decoder (GQA, q/k/v bias + M-RoPE in the qkv
    epilogue, LoRA K-extension) -> final RMSNorm on the last valid row -> optional SkipCA over the token-id-151643
    rows of hidden_states[0] -> value head. Not executed (output-identical): the reference's extra `self.visual(...)`
    pass (:356) and the 152064-wide lm_head."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._vplans: Dict[tuple, dict] = {}
        self.window_attn = "seg"   # "seg" (product) | "packed" (one tile per window; kept for comparison / tests)

    def rope_tables(self, n_pos: int, long: bool = False):
        """cos/sin [n_pos, head_dim/2] bf16 by position (Qwen2_5_VLRotaryEmbedding.forward, default rope,
        transformers modeling_qwen2_5_vl.py:595-608): fp32 angles -> cos/sin -> bf16."""
        key = (n_pos, False)
        if key not in self._rope:
            cfg = self.cfg
            expo = torch.arange(0, cfg.head_dim, 2, dtype=torch.int64, device=self.device).float() / cfg.head_dim
            inv_freq = 1.0 / (cfg.rope_theta ** expo)
            ang = torch.arange(n_pos, dtype=torch.int64, device=self.device).float()[:, None] * inv_freq[None, :]
            self._rope = {key: (ang.cos().to(torch.bfloat16).contiguous(), ang.sin().to(torch.bfloat16).contiguous())}
        return self._rope[key]

    def vision_plan(self, grids) -> dict:
        """Device copies of the host index plan for a tuple of (t, h, w) grids (cached: eval batches repeat shapes)."""
        key = tuple(grids)
        vp = self._vplans.get(key)
        if vp is None:
            cfg, dev = self.cfg, self.device
            from .weights import qwen_vit_padded_head_dim
            hp = qwen_window_plan(grids, cfg.vit_merge, cfg.vit_window, cfg.vit_patch)
            unit = cfg.vit_merge ** 2
            T = int(hp["src_row"].shape[0])
            # rotary rows per window-ordered token: [h-frequencies | w-frequencies] (rot_pos_emb, :382-409), fp32,
            # padded with cos = 1 / sin = 0 up to the padded head_dim / 2
            hd = cfg.vit_head_dim
            hdp = qwen_vit_padded_head_dim(hd)
            dim = hd // 2
            inv_freq = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float) / dim))
            pos = torch.from_numpy(hp["pos_hw"].astype(np.int64))
            freqs = torch.outer(torch.arange(int(pos.max()) + 1, dtype=torch.float), inv_freq)
            rot = freqs[pos].flatten(1)                                            # [T, hd/2]
            cos = torch.ones(T, hdp // 2, dtype=torch.float32)
            sin = torch.zeros(T, hdp // 2, dtype=torch.float32)
            cos[:, : hd // 2], sin[:, : hd // 2] = rot.cos(), rot.sin()
            # merger rows in the ORIGINAL order: merged unit g sits at window slot inv[g] (reverse_indices, :515-517)
            inv = np.argsort(hp["window_index"], kind="stable")
            unwin = (inv[:, None] * unit + np.arange(unit)[None, :]).reshape(-1).astype(np.int32)
            win_cu, img_cu = hp["win_cu"], hp["img_cu"]

            def d(a):
                return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

            wlen = np.diff(win_cu)
            row_lo = np.repeat(win_cu[:-1], wlen).astype(np.int32)     # per-token window bounds (lr_attention_seg_bf16)
            row_hi = np.repeat(win_cu[1:], wlen).astype(np.int32)
            vp = dict(T=T, src_row=d(hp["src_row"]), cos=cos.to(dev), sin=sin.to(dev), unwin=d(unwin),
                      row_lo=d(row_lo), row_hi=d(row_hi),
                      win_base=d(win_cu[:-1]), win_len=d(np.diff(win_cu).astype(np.int32)), n_win=len(win_cu) - 1,
                      win_max=int(np.diff(win_cu).max()),
                      img_base=d(img_cu[:-1]), img_len=d(np.diff(img_cu).astype(np.int32)), n_img=len(img_cu) - 1,
                      img_max=int(np.diff(img_cu).max()))
            if len(self._vplans) > 64:
                self._vplans.clear()
            self._vplans[key] = vp
        return vp

    def _vision_tower(self, pix: torch.Tensor, vp: dict) -> torch.Tensor:
        """-> merged image embeddings [T/4, H] in the original patch order."""
        from .weights import qwen_vit_padded_head_dim
        cfg, w = self.cfg, self.w
        T, D = vp["T"], cfg.vit_hidden
        nh, hd = cfg.vit_heads, cfg.vit_head_dim
        hdp = qwen_vit_padded_head_dim(hd)
        QW, AW = 3 * nh * hdp, nh * hdp
        K0, K0p = cfg.patch_dim, w.vit["patch_w"].shape[1]
        Ip = w.vit_layers[0]["dn_w"].shape[1] if w.vit_layers else 128
        a0 = self.buf("vit_a0", (T, K0p))
        ops.patch_rows(pix, vp["src_row"], a0, T, K0, K0p)
        x = self.buf("vit_x", (T, D))
        self._gemm(a0, w.vit["patch_w"], x, T, D, K0p)
        self._tap("vit_embed", x)
        hn = self.buf("vit_hn", (T, D))
        qkv = self.buf("vit_qkv", (T, QW))
        ao = self.buf("vit_ao", (T, AW))
        ff = self.buf("vit_ff", (T, Ip))
        scale = hd ** -0.5
        for li, lw in enumerate(w.vit_layers):
            ops.rmsnorm(x, lw["n1"], hn, T, D, cfg.vit_eps)
            ops.gemm_rope_ex(hn, lw["qkv_w"], qkv, T, QW, D, lw["qkv_b"], None, vp["cos"], vp["sin"], 2 * AW, hdp,
                             L.EPI_BIAS_ROPE_F32)
            if li in cfg.vit_fullatt:     # one packed sequence per image
                ops.attention_ex(qkv, qkv[:, AW:], qkv[:, 2 * AW:], ao, QW, AW, T, vp["n_img"], vp["img_max"],
                                 vp["img_base"], None, vp["img_len"], nh, nh, hdp, False, scale, L.ATTN_TCGEN05)
            elif self.window_attn == "seg":   # full 128-row tiles spanning several windows, per-row key ranges
                ops.attention_seg(qkv, qkv[:, AW:], qkv[:, 2 * AW:], ao, QW, AW, T, vp["row_lo"], vp["row_hi"], nh, hdp,
                                  scale)
            else:                         # one packed sequence (one 128-row tile) per window
                ops.attention_ex(qkv, qkv[:, AW:], qkv[:, 2 * AW:], ao, QW, AW, T, vp["n_win"], vp["win_max"],
                                 vp["win_base"], None, vp["win_len"], nh, nh, hdp, False, scale, L.ATTN_TCGEN05)
            self._gemm(ao, lw["proj_w"], x, T, D, AW, L.EPI_BIAS_RESIDUAL, lw["proj_b"], x)
            ops.rmsnorm(x, lw["n2"], hn, T, D, cfg.vit_eps)
            self._gemm(hn, lw["gu_w"], ff, T, 2 * Ip, D, L.EPI_BIAS_SWIGLU, lw["gu_b"])
            self._gemm(ff, lw["dn_w"], x, T, D, Ip, L.EPI_BIAS_RESIDUAL, lw["dn_b"], x)
            if li == 0:
                self._tap("vit_layer0", x)
        self._tap("vit_out", x)
        unit = cfg.vit_merge ** 2
        y = self.buf("vit_merge_in", (T, D))
        ops.rmsnorm(x, w.proj["ln_q"], y, T, D, 1e-6, row_index=vp["unwin"])
        Tm, D4, H = T // unit, D * unit, cfg.hidden_size
        y4 = y.view(Tm, D4)
        m1 = self.buf("vit_m1", (Tm, D4))
        self._gemm(y4, w.proj["m0_w"], m1, Tm, D4, D4, L.EPI_BIAS_GELU, w.proj["m0_b"])
        img = self.buf("img_proj", (Tm, H))
        self._gemm(m1, w.proj["m2_w"], img, Tm, H, D4, L.EPI_BIAS, w.proj["m2_b"])
        self._tap("image_embeds", img)
        return img

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, attention_mask: torch.Tensor, pixel_values: torch.Tensor,
                image_grid_thw, last_position: bool = False, mean_pool: bool = False) -> torch.Tensor:
        cfg, w, dev = self.cfg, self.w, self.device
        launches0 = L.launch_count()
        B, S = input_ids.shape
        H, M = cfg.hidden_size, B * S
        ids = input_ids.to(dev, torch.int64).contiguous()
        mask = attention_mask.to(dev, torch.int64).contiguous()
        pix = pixel_values.to(dev, torch.float32).contiguous()
        if pix.dim() != 2 or pix.shape[1] != cfg.patch_dim:
            raise ValueError(f"pixel_values of shape {tuple(pix.shape)}, expect [patches, {cfg.patch_dim}]")
        grid_dev = torch.as_tensor(image_grid_thw).to(dev, torch.int32).reshape(-1, 3).contiguous()
        n_images = grid_dev.shape[0]
        ca = cfg.add_cross_attention

        # 1. token plans (image positions; for SkipCA also the token-id-151643 positions) + M-RoPE plan, ONE D2H read
        pos = self.buf("pos", (M,), torch.int32)
        img_ord = self.buf("img_ord", (M,), torch.int32)
        nmeta = 4 * B + 1
        meta = self.buf("meta", (2 * nmeta + B + 3 * n_images,), torch.int32)
        seq_start, seq_len, eos_row, n_img, flags = (meta[:B], meta[B:2 * B], meta[2 * B:3 * B], meta[3 * B:4 * B],
                                                     meta[4 * B:nmeta])
        meta[4 * B:nmeta].zero_()
        ops.token_plan_ex(ids, mask, B, S, cfg.image_token_id, L.POS_ARANGE, pos, img_ord, seq_start, seq_len, eos_row,
                          n_img, flags)
        m2 = meta[nmeta:2 * nmeta]
        pad_ord = self.buf("pad_ord", (M,), torch.int32)
        if ca:
            m2[4 * B:].zero_()
            # the reference's mask is `input_ids == 151643` over ALL positions, padded ones included (:358)
            ops.token_plan_ex(ids, mask, B, S, cfg.pad_token_id, L.POS_ARANGE, pos, pad_ord, m2[:B], m2[B:2 * B],
                              m2[2 * B:3 * B], m2[3 * B:4 * B], m2[4 * B:])
        run_count = meta[2 * nmeta:2 * nmeta + B]
        meta[2 * nmeta + B:].copy_(grid_dev.reshape(-1))
        n_pos = S + 8
        cos_pos, sin_pos = self.rope_tables(n_pos)
        half = cfg.head_dim // 2
        pos3 = self.buf("pos3", (3, M), torch.int32)
        cos_tok = self.buf("cos_tok", (M, half))
        sin_tok = self.buf("sin_tok", (M, half))
        ops.mrope_plan(ids, mask, B, S, cfg.image_token_id, grid_dev, n_images, cfg.vit_merge, run_count, cos_pos, sin_pos,
                       n_pos, half, cfg.mrope_section[0], cfg.mrope_section[1], pos3, cos_tok, sin_tok, flags)
        meta_h = meta.cpu().numpy()
        if self.taps is not None:
            self.taps["pos3"] = pos3.clone()
        if meta_h[4 * B] & 1:
            raise ValueError("attention_mask rows must be one contiguous run of ones (left/right padding)")
        grids = [tuple(int(v) for v in meta_h[2 * nmeta + B + 3 * i: 2 * nmeta + B + 3 * i + 3]) for i in range(n_images)]
        unit = cfg.vit_merge ** 2
        n_tok = int(meta_h[3 * B:4 * B].sum())
        n_patch = sum(gt * gh * gw for gt, gh, gw in grids)
        n_feat = n_patch // unit
        if (meta_h[4 * B] & 2) or n_tok != n_feat:
            # get_placeholder_mask (transformers modeling_qwen2_5_vl.py:1204-1208) checks the batch total; a per-image
            # mismatch would shift features across images and is an error here
            raise ValueError(f"Image features and image tokens do not match, tokens: {n_tok}, features: {n_feat}")
        if n_patch != pix.shape[0]:
            raise ValueError(f"pixel_values has {pix.shape[0]} patches but image_grid_thw implies {n_patch}")
        for gt, gh, gw in grids:
            if gh % cfg.vit_merge or gw % cfg.vit_merge:
                raise ValueError(f"image grid {(gt, gh, gw)} is not a multiple of the merge size {cfg.vit_merge}")
        plan_h = np.zeros((2 * B, L.PLAN_STRIDE), dtype=np.int32)
        plan_h[:B, L.PLAN_NV] = meta_h[3 * B:4 * B]
        plan_h[:B, L.PLAN_ROW_BASE] = np.concatenate([[0], np.cumsum(meta_h[3 * B:4 * B])[:-1]])
        n_pad = meta_h[nmeta + 3 * B: nmeta + 4 * B] if ca else np.zeros(B, dtype=np.int32)
        plan_h[B:, L.PLAN_NV] = n_pad
        plan_h[B:, L.PLAN_ROW_BASE] = np.concatenate([[0], np.cumsum(n_pad)[:-1]])
        sum_pad, max_pad = int(n_pad.sum()), int(n_pad.max()) if B else 0
        host = torch.from_numpy(plan_h.reshape(-1))
        dev_plan = self.buf("plan", (host.numel(),), torch.int32)
        self._h2d("plan", host, dev_plan)
        plan, pad_plan = dev_plan[: B * L.PLAN_STRIDE], dev_plan[B * L.PLAN_STRIDE:]

        # 2. vision tower + merger, 3. embeddings
        vp = self.vision_plan(grids)
        img = self._vision_tower(pix, vp)
        hid = self.buf("hidden", (M, H))
        ops.embed_scatter(ids, img_ord, plan, w.embed, img, hid, B, S, H, cfg.vocab_size)
        self._tap("inputs_embeds", hid)
        kv_src = None
        if ca and sum_pad > 0:
            kv_src = self.buf("ca_src", (sum_pad, H))
            ops.compact_rows(hid, pad_ord, pad_plan, kv_src, B, S, H)

        # 4. decoder (per-token M-RoPE rows: position_ids = None)
        head_row = self._head_rows(eos_row, B, S, last_position)
        if self._can_pack(meta_h, B, S, last_position, mean_pool):
            # per-token M-RoPE rows travel with their tokens: the same gather on the cos / sin tables
            hid, _, base_p, head_row, rows_p, longest, idx_p = self._pack_valid_rows(hid, meta_h, B, S, True)
            cos_p = self.buf("cos_tok_packed", (M, half))[:rows_p]
            sin_p = self.buf("sin_tok_packed", (M, half))[:rows_p]
            ops.gather_rows(cos_tok, idx_p, cos_p, rows_p, half)
            ops.gather_rows(sin_tok, idx_p, sin_p, rows_p, half)
            hid_e = self._decoder(hid, B, S, None, None, seq_len, cos_p, sin_p, head_row,
                                  packed=(base_p, rows_p, longest))
        else:
            hid_e = self._decoder(hid, B, S, None, seq_start, seq_len, cos_tok, sin_tok,
                                  None if mean_pool else head_row)
        if mean_pool:
            if ca:   # all-rows form of the qwen SkipCA arm on the compacted token-id-151643 rows of hidden_states[0]
                reward = self._mean_head(hid, mask, B, S, True, kv_src, plan_h[B:], max_pad if sum_pad > 0 else 0,
                                         cross_attention=True, masked_pad=True, ca_eps=1e-6)
            else:
                reward = self._mean_head(hid, mask, B, S, True, cross_attention=False)
            self.launches = L.launch_count() - launches0
            return reward

        # 5. final norm on the last valid row, SkipCA (qwen arm), value head
        xe = self._final_rows(hid, hid_e, head_row, B)
        vhd = cfg.vhd
        reward = torch.empty(B, vhd, dtype=torch.bfloat16, device=dev)
        if ca and sum_pad > 0:
            q = self.buf("ca_q", (B, H))
            self._gemm(xe, w.head["wq"], q, B, H, H)
            kv = self.buf("ca_kv", (sum_pad, 2 * H))
            self._gemm(kv_src, w.head["wkv"], kv, sum_pad, 2 * H, H)
            scores = self.buf("ca_scores", (B, max_pad), torch.float32)
            ops.skipca_scores_ex(q, kv, pad_plan, scores, B, H, max_pad, -9984.0)   # bf16(-1e4), masked_fill at :391
            ops.skipca_head(scores, kv, pad_plan, xe, w.head["ca_ln"], w.head["vh"], reward, B, H, max_pad, vhd, 1e-6)
        elif ca:
            # no token-id-151643 position in the batch: vision_pad is [B, 0, H], attn_o = 0, ca_layernorm(last + 0)
            xc = self.buf("x_ca", (max(B, 1), H))
            ops.rmsnorm(xe, w.head["ca_ln"], xc, B, H, 1e-6)
            ops.skipca_head(None, None, None, xc, None, w.head["vh"], reward, B, H, 0, vhd, cfg.rms_eps)
        else:
            ops.skipca_head(None, None, None, xe, None, w.head["vh"], reward, B, H, 0, vhd, cfg.rms_eps)
        self.launches = L.launch_count() - launches0
        return reward
