"""B200-native drop-in for the reference's missing `pretrain_src/model/pretrain_goat.py`.

`GlocalTextPathCMTPreTraining` keeps the reference module API:
  * `from_pretrained(pretrained_model_name_or_path=None, config=..., state_dict=...)`
    (call sites pretrain_src/train_r2r_magic.py:260-262, :275-277; non-strict, tolerates foreign keys)
  * `forward(batch, task, compute_loss=True)` with tasks mlm / sap / mrc / cfp
    (outputs pinned at train_r2r_magic.py:448, :483, :510-512, :545-546)
  * the parameter tree `bert.embeddings.*`, `bert.lang_encoder.layer.{i}.*`,
    `bert.{local,global}_encoder.encoder.crossattention.{i}.*` (train_r2r_magic.py:189-208)
  * `torch.nn.Dropout` sub-modules whose `p` the driver rewrites (pretrain_src/utils/misc.py:19-25).

The nn.Module tree below only HOLDS parameters; all math runs in the hand-written sm_100a kernels of
libmagic_b200 through ops.py.  There is no PyTorch fallback: tensors must live on a CUDA device.
Architecture decisions are those frozen in SURVEY.md Appendix A.
"""
import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .graph_index import INDEX_KEY, build_index, index_to


def _cfg(c, name, default):
    return getattr(c, name, default)


# ---------------------------------------------------------------------------------------------------
# parameter containers (names mirror HF BertLayer / METER BertCrossLayer / nn.TransformerEncoderLayer)
# ---------------------------------------------------------------------------------------------------
class _SelfAtt(nn.Module):
    def __init__(self, c):
        super().__init__()
        h = c.hidden_size
        self.query, self.key, self.value = nn.Linear(h, h), nn.Linear(h, h), nn.Linear(h, h)
        self.dropout = nn.Dropout(c.attention_probs_dropout_prob)


class _SelfOut(nn.Module):
    def __init__(self, c, in_size=None):
        super().__init__()
        self.dense = nn.Linear(in_size or c.hidden_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.dropout = nn.Dropout(c.hidden_dropout_prob)


class _Attention(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.self = _SelfAtt(c)
        self.output = _SelfOut(c)


class _Intermediate(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.intermediate_size)


class _BertLayer(nn.Module):
    def __init__(self, c, cross=False):
        super().__init__()
        self.attention = _Attention(c)
        if cross:
            self.crossattention = _Attention(c)
        self.intermediate = _Intermediate(c)
        self.output = _SelfOut(c, c.intermediate_size)


class _LangEncoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.layer = nn.ModuleList([_BertLayer(c) for _ in range(c.num_l_layers)])


class _CrossEncoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.crossattention = nn.ModuleList([_BertLayer(c, cross=True) for _ in range(c.num_x_layers)])


class _Embeddings(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.word_embeddings = nn.Embedding(c.vocab_size, c.hidden_size)
        self.position_embeddings = nn.Embedding(c.max_position_embeddings, c.hidden_size)
        self.token_type_embeddings = nn.Embedding(c.type_vocab_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.dropout = nn.Dropout(c.hidden_dropout_prob)


class _MHA(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * h, h))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * h))
        self.out_proj = nn.Linear(h, h)


class _PanoLayer(nn.Module):
    def __init__(self, c):
        super().__init__()
        h = c.hidden_size
        self.self_attn = _MHA(h)
        self.linear1 = nn.Linear(h, c.intermediate_size)
        self.linear2 = nn.Linear(c.intermediate_size, h)
        self.norm1 = nn.LayerNorm(h, eps=c.layer_norm_eps)
        self.norm2 = nn.LayerNorm(h, eps=c.layer_norm_eps)
        self.dropout = nn.Dropout(c.hidden_dropout_prob)
        self.dropout1 = nn.Dropout(c.hidden_dropout_prob)
        self.dropout2 = nn.Dropout(c.hidden_dropout_prob)


class _PanoEncoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.layers = nn.ModuleList([_PanoLayer(c) for _ in range(c.num_pano_layers)])
        self.norm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)


class _ImageEmbeddings(nn.Module):
    def __init__(self, c):
        super().__init__()
        h = c.hidden_size
        self.img_linear = nn.Linear(c.image_feat_size, h)
        self.img_layer_norm = nn.LayerNorm(h, eps=c.layer_norm_eps)
        self.loc_linear = nn.Linear(c.angle_feat_size + 3, h)
        self.loc_layer_norm = nn.LayerNorm(h, eps=c.layer_norm_eps)
        self.nav_type_embedding = nn.Embedding(3, h)
        if _cfg(c, "obj_feat_size", 0) > 0:  # REVERIE / SOON object tokens (DUET lineage names)
            self.obj_linear = nn.Linear(c.obj_feat_size, h)
            self.obj_layer_norm = nn.LayerNorm(h, eps=c.layer_norm_eps)
        else:
            self.obj_linear = self.obj_layer_norm = None
        self.layer_norm = nn.LayerNorm(h, eps=c.layer_norm_eps)
        self.dropout = nn.Dropout(c.hidden_dropout_prob)
        self.pano_encoder = _PanoEncoder(c) if c.num_pano_layers > 0 else None
        self.adaptive_pano_attn = nn.Linear(h, 1) if c.adaptive_pano_fusion else None


class _LocalVPEncoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.vp_pos_embeddings = nn.Sequential(
            nn.Linear(c.angle_feat_size * 2 + 6, c.hidden_size), nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps))
        self.encoder = _CrossEncoder(c)


class _GlobalMapEncoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.gmap_pos_embeddings = nn.Sequential(
            nn.Linear(c.angle_feat_size + 3, c.hidden_size), nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps))
        self.gmap_step_embeddings = nn.Embedding(c.max_action_steps, c.hidden_size)
        self.encoder = _CrossEncoder(c)
        self.sprel_linear = nn.Linear(1, 1) if c.graph_sprels else None


class _ClsPrediction(nn.Module):
    def __init__(self, h, input_size=None, out=1, eps=1e-12):
        super().__init__()
        input_size = h if input_size is None else input_size
        self.net = nn.Sequential(nn.Linear(input_size, h), nn.ReLU(), nn.LayerNorm(h, eps=eps), nn.Linear(h, out))


class _Transform(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)


class _LMPredictionHead(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.transform = _Transform(c)
        self.decoder = nn.Linear(c.hidden_size, c.vocab_size, bias=False)
        self.bias = nn.Parameter(torch.zeros(c.vocab_size))


class _MLMHead(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.predictions = _LMPredictionHead(c)


KD_HEADS = ("txt_emb_w", "kdl_img_w", "kdl_avg_img_w", "global_cross_w", "local_cross_w", "vp_txt_w", "gmap_txt_w")
# learned ability weights, used through softplus (agent.py:585,618,678,681,713); order = the RW weight order
KD_LEARNED_WEIGHTS = ("kdl_txt_weight", "kdl_img_weight", "kdl_global_weight", "kdl_local_weight", "kdl_predict_weight")
SOFTPLUS_ONE = 0.5413248546129181  # softplus(x) = 1: a learned ability weight starts at the non-adaptive value


def _kdl_get(c, key, default=None):
    kdl = getattr(c, "kdl", None)
    if kdl is None:
        return default
    return kdl.get(key, default) if hasattr(kdl, "get") else getattr(kdl, key, default)


# ---------------------------------------------------------------------------------------------------
# functional forward built from the fused kernels
# ---------------------------------------------------------------------------------------------------
class _Ctx:
    """Per-forward state: compute dtype, training flag, dropout salts, KD attention maps on/off."""

    def __init__(self, dtype, training, want_attn, salt_base):
        self.dtype, self.training, self.want_attn = dtype, training, want_attn
        self._salt = salt_base

    def salt(self):
        self._salt += 1
        return self._salt

    def p(self, dropout_module):
        return float(dropout_module.p) if self.training else 0.0


def _ln(x, m, fc, res=None, p_in=0.0, p_out=0.0, link=None):
    return ops.layer_norm(x, m.weight, m.bias, m.eps, res=res, p_in=p_in, salt_in=fc.salt(), p_out=p_out,
                          salt_out=fc.salt(), link=link)


def _attn_block(att, x, ctx_kv, B, Lq, Lk, key_lens, fc, dists=None, sprel=None, key_skip=-1, want_map=True):
    """BertAttention: (self or cross) attention + output dense + dropout + residual LayerNorm.
    x [B*Lq, h]; ctx_kv None -> self attention.  Returns (y [B*Lq, h], pbar or None)."""
    sa = att.self
    h = x.shape[-1]
    H = h // 64
    pdrop = fc.p(sa.dropout)
    sw = sprel.weight if sprel is not None else None
    sb = sprel.bias if sprel is not None else None
    # y = LN(x + f(x)): the residual-branch gradient rides the dgrad GEMM of f's first linear (ops.GradLink)
    link = ops.GradLink() if (x.requires_grad and torch.is_grad_enabled()) else None
    if ctx_kv is None:
        qkv = ops.packed_linear(x, [sa.query.weight, sa.key.weight, sa.value.weight],
                                [sa.query.bias, sa.key.bias, sa.value.bias], link=link)
        o, pbar = ops.attention(qkv, None, 0, h, 2 * h, B, H, Lq, Lk, key_lens, dists, sw, sb,
                                fc.want_attn and want_map, pdrop, fc.salt(), key_skip)
    else:
        q = ops.linear(x, sa.query.weight, sa.query.bias, link=link)
        kv = ops.packed_linear(ctx_kv, [sa.key.weight, sa.value.weight], [sa.key.bias, sa.value.bias])
        o, pbar = ops.attention(q, kv, 0, 0, h, B, H, Lq, Lk, key_lens, None, None, None, fc.want_attn and want_map,
                                pdrop, fc.salt())
    so = att.output
    d = ops.linear(o, so.dense.weight, so.dense.bias)
    y = _ln(d, so.LayerNorm, fc, res=x, p_in=fc.p(so.dropout), link=link)
    return y, pbar


def _ffn_block(layer, x, fc):
    link = ops.GradLink() if (x.requires_grad and torch.is_grad_enabled()) else None
    f = ops.ffn(x, layer.intermediate.dense.weight, layer.intermediate.dense.bias, layer.output.dense.weight,
                layer.output.dense.bias, act=L.ACT_GELU, link=link)
    return _ln(f, layer.output.LayerNorm, fc, res=x, p_in=fc.p(layer.output.dropout), link=link)


def _cross_encoder(enc, x, ctx, B, Lx, Lc, x_lens, c_lens, fc, dists=None, sprel=None, key_skip=-1, map_depth=None):
    """`map_depth`: KD attention maps are only produced for the first `map_depth` layers (None = all): MAKD compares
    the maps of teacher and student on their common depth (agent.py:560,654,671), deeper maps are never read."""
    attns = []
    for li, layer in enumerate(enc.crossattention):
        wm = map_depth is None or li < map_depth
        a, p_self = _attn_block(layer.attention, x, None, B, Lx, Lx, x_lens, fc, dists, sprel, key_skip, want_map=wm)
        cx, p_cross = _attn_block(layer.crossattention, a, ctx, B, Lx, Lc, c_lens, fc, want_map=wm)
        x = _ffn_block(layer, cx, fc)
        attns.append((p_self, p_cross))
    return x, attns


class GlocalTextPathCMT(nn.Module):
    # data-parallel training: an object with expect(stage, key) / fire(stage, key) (parallel.StageSync).  The forward
    # marks the activations at which backward has FINISHED a group of layers; the tensor hook fires when that
    # activation's gradient is complete, i.e. every layer above it has issued its weight gradients, so the gradient
    # arena range of those layers can be exchanged while the layers below still back-propagate.
    stage_cb = None
    # (text layers, cross-modal layers) whose KD attention maps are needed; None = all.  MAKD compares the maps of the two
    # models on their common depth only (agent.py:560), so the deeper model skips the rest (train_step sets this)
    kd_attn_depth = None

    def _mark(self, t, stage, key):
        cb = self.stage_cb
        if cb is not None and torch.is_grad_enabled() and t.requires_grad:
            cb.expect(stage, key)
            t.register_hook(lambda g, s=stage, k=key: cb.fire(s, k))
        return t

    def __init__(self, c):
        super().__init__()
        self.config = c
        self.embeddings = _Embeddings(c)
        self.lang_encoder = _LangEncoder(c)
        self.img_embeddings = _ImageEmbeddings(c)
        self.local_encoder = _LocalVPEncoder(c)
        self.global_encoder = _GlobalMapEncoder(c)
        ht = _cfg(c, "teacher_hidden_size", None)
        if _cfg(c, "role", "student") == "student" and _cfg(c, "kd", False) and ht:
            for name in KD_HEADS:
                setattr(self, name, nn.Linear(c.hidden_size, ht))
            if _kdl_get(c, "kdl_adaptive_ability_weight_type") == "learned_weight":
                for name in KD_LEARNED_WEIGHTS:  # [DECISION] one scalar each, initialised so that softplus = 1
                    setattr(self, name, nn.Parameter(torch.full((1,), SOFTPLUS_ONE)))

    # -- text ------------------------------------------------------------------------------------
    def forward_text(self, batch, ix, fc):
        e = self.embeddings
        ids = batch["txt_ids"]
        B, Lt = ids.shape
        typ = e.token_type_embeddings.weight if e.token_type_embeddings.weight.shape[0] == 1 else \
            e.token_type_embeddings.weight[0].contiguous()
        x = ops.embed_ln(ids, e.word_embeddings.weight, e.position_embeddings.weight, typ, e.LayerNorm.weight,
                         e.LayerNorm.bias, e.LayerNorm.eps, fc.dtype, fc.p(e.dropout), fc.salt())
        x = x.view(B * Lt, -1)
        attns = []
        from .parallel import text_stage_cuts
        cuts = text_stage_cuts(len(self.lang_encoder.layer)) if self.stage_cb is not None else []
        for li, layer in enumerate(self.lang_encoder.layer):
            if li in cuts:  # gradient complete <=> text layers li.. have finished backward (stage 1 + position in cuts)
                self._mark(x, 1 + cuts.index(li), f"txt_{li}")
            a, p = _attn_block(layer.attention, x, None, B, Lt, Lt, ix["key_lens_txt"], fc,
                               want_map=self.kd_attn_depth is None or li < self.kd_attn_depth[0])
            x = _ffn_block(layer, a, fc)
            attns.append(p)
        if not _cfg(self.config, "update_lang_bert", True):
            x = x.detach()
        return x, attns

    # -- panorama ----------------------------------------------------------------------------------
    def forward_pano(self, batch, ix, fc, img_fts=None):
        ie = self.img_embeddings
        e = self.embeddings
        if img_fts is not None:
            fts = img_fts
        elif batch.get("traj_view_img_fts") is not None:
            fts = batch["traj_view_img_fts"]
        else:
            fts = self.view_features(batch, fc.dtype)
        R, V, Fd = fts.shape
        h = self.config.hidden_size
        x_in = fts.reshape(R * V, Fd)
        if x_in.dtype != fc.dtype:
            # materialised fp32 features (the reference collate's layout) in bf16 mode: one cast kernel.  The feature
            # store (featurizer.FeatureStore, `view_features` above) gathers straight into the compute dtype instead
            x_in = x_in.to(fc.dtype)
        z = ops.linear(x_in, ie.img_linear.weight, ie.img_linear.bias)
        a = _ln(z, ie.img_layer_norm, fc)
        view_lens = batch["traj_vp_view_lens"]
        if batch.get("traj_obj_img_fts") is not None and ie.obj_linear is not None:
            # object tokens follow each panorama's views ([cand_views, noncand_views, objs], dataset.py:447): embed
            # them with their own linear + LN and interleave by the host-built row tables (-1 = padding -> zero row)
            ofts = batch["traj_obj_img_fts"]
            o_in = ofts.reshape(-1, ofts.shape[-1])
            if o_in.dtype != fc.dtype:
                o_in = o_in.to(fc.dtype)
            ao = _ln(ops.linear(o_in, ie.obj_linear.weight, ie.obj_linear.bias), ie.obj_layer_norm, fc)
            a = ops.add(ops.gather_rows(a, ix["pano_view_src"]), ops.gather_rows(ao, ix["pano_obj_src"]))
            V = batch["traj_loc_fts"].shape[1]
            view_lens = ix["pano_lens"]
        typ = e.token_type_embeddings.weight if e.token_type_embeddings.weight.shape[0] == 1 else \
            e.token_type_embeddings.weight[0].contiguous()
        s = ops.posfuse(a, batch["traj_nav_types"].reshape(-1), ie.nav_type_embedding.weight, typ,
                        batch["traj_loc_fts"].reshape(R * V, -1), ie.loc_linear.weight, ie.loc_linear.bias,
                        ie.loc_layer_norm.weight, ie.loc_layer_norm.bias, ie.loc_layer_norm.eps, fc.dtype)
        x = _ln(s, ie.layer_norm, fc, p_out=fc.p(ie.dropout))
        attns = []
        H = h // 64
        if ie.pano_encoder is not None:
            if self.stage_cb is not None:  # gradient complete <=> the panorama encoder has finished backward (last stage)
                from .parallel import text_stage_cuts
                self._mark(x, 2 + len(text_stage_cuts(len(self.lang_encoder.layer))) - 0, "pano_in")
            for layer in ie.pano_encoder.layers:
                n1 = _ln(x, layer.norm1, fc)
                mha = layer.self_attn
                qkv = ops.linear(n1, mha.in_proj_weight, mha.in_proj_bias)
                o, p = ops.attention(qkv, None, 0, h, 2 * h, R, H, V, V, ix["key_lens_pano"], None, None, None,
                                     fc.want_attn, fc.p(layer.dropout), fc.salt())
                x = ops.linear(o, mha.out_proj.weight, mha.out_proj.bias, residual=x, drop_p=fc.p(layer.dropout1),
                               salt=fc.salt())
                n2 = _ln(x, layer.norm2, fc)
                x = ops.ffn(n2, layer.linear1.weight, layer.linear1.bias, layer.linear2.weight, layer.linear2.bias,
                            act=L.ACT_GELU, residual=x, drop_p=fc.p(layer.dropout), salt=fc.salt(),
                            drop_out_p=fc.p(layer.dropout2), salt_out=fc.salt())
                attns.append(p)
            x = _ln(x, ie.pano_encoder.norm, fc)
        x3 = x.view(R, V, h)
        ap = ie.adaptive_pano_attn
        fused = ops.pano_fuse(x3, ap.weight if ap is not None else None, ap.bias if ap is not None else None,
                              view_lens)
        return x3, fused, attns

    feature_store = None

    def view_features(self, batch, dtype):
        """Compact batches (featurizer.py) carry panorama rows + view orders instead of 36 x 768 features per step: the
        features are gathered from the device-resident store, directly in the compute dtype."""
        store = getattr(self, "feature_store", None)
        if store is None or batch.get("traj_vp_index") is None:
            raise ValueError("the batch has no traj_view_img_fts and no feature store is attached "
                             "(featurizer.FeatureStore.attach)")
        return store.gather_views(batch["traj_vp_index"], batch["traj_view_perm"], out_dtype=dtype)

    def gmap_input(self, pano, fused, batch, ix, fc):
        ge = self.global_encoder
        B, G = batch["gmap_step_ids"].shape
        feats = ops.gmap_aggregate(pano, fused, ix)
        pe = ge.gmap_pos_embeddings
        return ops.posfuse(feats, batch["gmap_step_ids"].reshape(-1), ge.gmap_step_embeddings.weight, None,
                           batch["gmap_pos_fts"].reshape(B * G, -1), pe[0].weight, pe[0].bias, pe[1].weight,
                           pe[1].bias, pe[1].eps, fc.dtype)

    def vp_input(self, pano, batch, ix, fc):
        B, Vp = batch["vp_pos_fts"].shape[:2]
        x_in = ops.gather_rows(pano, ix["vp_gather"])
        pe = self.local_encoder.vp_pos_embeddings
        return ops.posfuse(x_in, None, None, None, batch["vp_pos_fts"].reshape(B * Vp, -1), pe[0].weight, pe[0].bias,
                           pe[1].weight, pe[1].bias, pe[1].eps, fc.dtype)

    def forward(self, batch, mode, fc, ix, img_fts=None):
        """The text and panorama encoders are independent, and so are the global and local cross-modal encoders:
        each pair runs as two concurrent stream branches (every kernel here fills only a fraction of the 148 SMs,
        so the step is bound by the length of the dependent-kernel chain, not by throughput).  Autograd replays
        each branch's backward on the stream its forward ran on; under CUDA-graph capture the branches become
        parallel paths of the graph."""
        B, Lt = batch["txt_ids"].shape
        G = batch["gmap_step_ids"].shape[1]
        Vp = batch["vp_pos_fts"].shape[1]
        h = self.config.hidden_size
        ge, le = self.global_encoder, self.local_encoder

        def visual():
            pano, fused, img_attns = self.forward_pano(batch, ix, fc, img_fts)
            return pano, fused, img_attns, self.gmap_input(pano, fused, batch, ix, fc), self.vp_input(pano, batch, ix, fc)

        (pano, fused, img_attns, g_in, v_in), (txt, txt_attns) = ops.run_branches(
            visual, lambda: self.forward_text(batch, ix, fc))
        # stage 0 of the gradient exchange: the cross-modal encoders and every head have finished backward once the
        # gradients of their three inputs are complete
        for t_, k_ in ((txt, "txt"), (g_in, "g_in"), (v_in, "v_in")):
            self._mark(t_, 0, k_)
        xd = None if self.kd_attn_depth is None else self.kd_attn_depth[1]
        if mode == "nav":
            dists = batch["gmap_pair_dists"] if ge.sprel_linear is not None else None
            (v, v_attn), (g, g_attn) = ops.run_branches(
                lambda: _cross_encoder(le.encoder, v_in, txt, B, Vp, Lt, ix["key_lens_vp"], ix["key_lens_txt"], fc,
                                       map_depth=xd),
                lambda: _cross_encoder(ge.encoder, g_in, txt, B, G, Lt, ix["key_lens_gmap"], ix["key_lens_txt"], fc,
                                       dists, ge.sprel_linear, map_depth=xd))
            g, v = g.view(B, G, h), v.view(B, Vp, h)
        else:
            (v, v_attn), (g, g_attn) = ops.run_branches(
                lambda: _cross_encoder(le.encoder, txt, v_in, B, Lt, Vp, ix["key_lens_txt"], ix["key_lens_vp"], fc,
                                       map_depth=xd),
                lambda: _cross_encoder(ge.encoder, txt, g_in, B, Lt, G, ix["key_lens_txt"], ix["key_lens_gmap"], fc,
                                       map_depth=xd))
            g, v = g.view(B, Lt, h), v.view(B, Lt, h)
        return dict(txt_embeds=txt.view(B, Lt, h), txt_attn_list=txt_attns, pano_embeds=pano,
                    pano_fused_embeds=fused, img_attn_list=img_attns, gmap_embeds=g, gmap_attn_list=g_attn,
                    vp_embeds=v, vp_attn_list=v_attn, pano_row_scale=ix.get("pano_row_scale"))


_TASK_HEADS = {"mlm": ("mlm_head.",), "sap": ("global_sap_head.", "local_sap_head.", "sap_fuse_linear."),
               "mrc": ("image_classifier.",), "og": ("og_head.",),
               "cfp": ("cfp_gmap_proj.", "cfp_vp_proj.", "cfp_txt_proj.")}
_ALL_HEADS = tuple(p for ps in _TASK_HEADS.values() for p in ps)
_KD_COMMON = ("bert.txt_emb_w.", "bert.kdl_img_w.", "bert.kdl_avg_img_w.")
_KD_MLM = ("bert.vp_txt_w.", "bert.gmap_txt_w.")
_KD_NAV = ("bert.global_cross_w.", "bert.local_cross_w.")


def inactive_in_task(task, name, kd=False):
    """True when parameter `name` of GlocalTextPathCMTPreTraining receives NO gradient in a step of `task` (read off
    the forward functions above): the heads of the other tasks, `sprel_linear` in the MLM branch (text queries attend
    the graph without the distance bias), and the KD projections the step's MAKD losses do not use (all of them
    without a teacher).  The reference optimizer skips exactly these (`p.grad is None`, optim/adamw.py:66-67)."""
    t = task[:3] if task[:3] in _TASK_HEADS else task
    if name.startswith(_ALL_HEADS):
        return not name.startswith(_TASK_HEADS.get(t, ()))
    if name.startswith("bert.global_encoder."):
        if t in ("mrc", "og"):  # these heads read the LOCAL branch only: the global branch gets no gradient
            return True
        return t == "mlm" and name.startswith("bert.global_encoder.sprel_linear.")
    if name.startswith(_KD_COMMON) or name.startswith("bert.kdl_") and name.endswith("_weight"):
        return not (kd and t in ("mlm", "sap"))
    if name.startswith(_KD_MLM):
        return not (kd and t == "mlm")
    if name.startswith(_KD_NAV):
        return not (kd and t == "sap")
    if name.startswith("bert.img_embeddings.adaptive_pano_attn."):
        return t in ("mrc", "og")  # the fused panorama embedding only feeds the global branch's visited nodes
    if name.startswith("bert.img_embeddings.obj_"):
        return t != "og"
    return False


def stack_attns(lst):
    """KD attention maps in the oracle's 4-D layout: [B, n_layers, Lq, Lk] ([self | cross] for x-layers)."""
    lst = [a for a in lst if (a[0] if isinstance(a, tuple) else a) is not None]  # (maps beyond kd_attn_depth)
    if not lst:
        return None
    if isinstance(lst[0], tuple):
        return torch.stack([torch.cat([a, b], -1) for a, b in lst], 1)
    return torch.stack(lst, 1)


class GlocalTextPathCMTPreTraining(nn.Module):
    def __init__(self, config):
        super().__init__()
        c = self.config = config
        for k, v in dict(layer_norm_eps=1e-12, type_vocab_size=1, max_action_steps=100, image_feat_size=768,
                         image_prob_size=1000, angle_feat_size=4, graph_sprels=True, glocal_fuse=True,
                         adaptive_pano_fusion=True, cfp_temperature=1.0, initializer_range=0.02,
                         max_position_embeddings=514, vocab_size=50265, hidden_dropout_prob=0.1,
                         attention_probs_dropout_prob=0.1, num_pano_layers=2).items():
            if not hasattr(c, k):
                setattr(c, k, v)
        if c.hidden_size % 64 != 0:
            raise ValueError("hidden_size must be a multiple of 64 (heads = hidden/64, train_r2r_magic.py:143,157)")
        self.bert = GlocalTextPathCMT(c)
        tasks = set(_cfg(c, "pretrain_tasks", ("mlm", "sap")))
        h = c.hidden_size
        if "mlm" in tasks:
            self.mlm_head = _MLMHead(c)
            self.mlm_head.predictions.decoder.weight = self.bert.embeddings.word_embeddings.weight  # tied
        if "sap" in tasks or "cfp" in tasks:
            self.global_sap_head = _ClsPrediction(h, eps=c.layer_norm_eps)
            self.local_sap_head = _ClsPrediction(h, eps=c.layer_norm_eps)
            self.sap_fuse_linear = _ClsPrediction(h, input_size=2 * h, eps=c.layer_norm_eps) if c.glocal_fuse else None
        if "mrc" in tasks:
            self.image_classifier = _ClsPrediction(h, out=c.image_prob_size, eps=c.layer_norm_eps)
        if "cfp" in tasks:
            for n in ("cfp_gmap_proj", "cfp_vp_proj", "cfp_txt_proj"):
                setattr(self, n, nn.Linear(h, h))
        if "og" in tasks:
            if _cfg(c, "obj_feat_size", 0) <= 0:
                raise ValueError("the og task needs obj_feat_size > 0 (r2r_magic_model_config.json:48 sets 0: R2R / RxR "
                                 "have no object features)")
            self.og_head = _ClsPrediction(h, eps=c.layer_norm_eps)
        self.compute_dtype = torch.float32
        self.output_kd = bool(_cfg(c, "kd", False))
        self.feature_store = None  # featurizer.FeatureStore.attach(model)
        self.apply(self._init_weights)

    def __getattr__(self, name):
        """The fine-tune agent reaches the KD heads and the learned ability weights directly on the model it wraps
        (`s_model = self.vln_bert.vln_bert; s_model.txt_emb_w(...)`, `s_model.kdl_txt_weight`: agent.py:552-568,585);
        here they live on `.bert` with the encoders, so those names -- and only those -- resolve through it."""
        try:
            return super().__getattr__(name)
        except AttributeError:
            if name in KD_HEADS or name in KD_LEARNED_WEIGHTS:
                return getattr(super().__getattr__("bert"), name)
            raise

    # -- construction ------------------------------------------------------------------------------
    def _init_weights(self, m):
        r = self.config.initializer_range
        if isinstance(m, (nn.Linear, nn.Embedding)):
            m.weight.data.normal_(0.0, r)
        elif isinstance(m, nn.LayerNorm):
            m.weight.data.fill_(1.0)
            m.bias.data.zero_()
        elif isinstance(m, _MHA):
            m.in_proj_weight.data.normal_(0.0, r)
            m.in_proj_bias.data.zero_()
        if isinstance(m, nn.Linear) and m.bias is not None:
            m.bias.data.zero_()

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path=None, config=None, state_dict=None, **kwargs):
        """Non-strict load, like transformers' `PreTrainedModel.from_pretrained(None, config=, state_dict=)`
        as used at train_r2r_magic.py:260-262: unknown keys are ignored, missing keys keep their init, a
        `module.` prefix is stripped, values may be nn.Parameter."""
        if config is None:
            raise ValueError("config is required")
        model = cls(config)
        if state_dict:
            own = model.state_dict()
            loaded = {}
            for k, v in state_dict.items():
                k2 = k[7:] if k.startswith("module.") else k
                if k2 in own and tuple(own[k2].shape) == tuple(v.shape):
                    loaded[k2] = v.detach() if isinstance(v, torch.Tensor) else torch.as_tensor(v)
            model.load_state_dict(loaded, strict=False)
            model.load_info = dict(loaded=sorted(loaded), missing=sorted(set(own) - set(loaded)),
                                   unexpected=sorted(set(k for k in state_dict) - set(loaded)))
        return model

    def set_kd_attn_depth(self, n_text=None, n_cross=None):
        """Only the first `n_text` text layers / `n_cross` cross-modal layers produce KD attention maps (None: all)."""
        self.bert.kd_attn_depth = None if n_text is None else (int(n_text), int(n_cross if n_cross is not None else n_text))
        return self

    def set_compute_dtype(self, dtype):
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("compute dtype must be float32 or bfloat16")
        self.compute_dtype = dtype
        return self

    # -- forward -------------------------------------------------------------------------------------
    def _fc(self, want_attn):
        role_base = 0 if _cfg(self.config, "role", "student") == "student" else 1 << 20
        return _Ctx(self.compute_dtype, self.training, want_attn, role_base)

    def _index(self, batch):
        dev = batch["txt_ids"].device
        if not dev.type == "cuda":
            raise L.MagicError("magic_b200 runs on CUDA tensors only (no CPU fallback); move the batch to the GPU")
        ix = batch.get(INDEX_KEY)
        if ix is None:
            ix = build_index(batch)  # NOTE: syncs (reads masks back); prefer graph_index.prepare_batch() on the host
        if ix["node_ptr"].device != dev:
            ix = index_to(ix, dev)
            batch[INDEX_KEY] = ix
        return ix

    def forward(self, batch, task, compute_loss=True, output_kd=None):
        """`output_kd` (ours): also return the KD attention maps; default = config.kd (train_r2r_magic.py:134,159)."""
        arena = getattr(self, "_magic_arena", None)
        if arena is not None and not torch.cuda.is_current_stream_capturing():
            arena.sync_lowp()  # a checkpoint loaded after the arena was built must reach the bf16 GEMM operands
        if output_kd is not None:
            prev, self.output_kd = self.output_kd, bool(output_kd)
            try:
                return self.forward(batch, task, compute_loss)
            finally:
                self.output_kd = prev
        if task.startswith("mlm"):
            return self.forward_mlm(batch, compute_loss)
        if task.startswith("sap"):
            return self.forward_sap(batch, compute_loss)
        if task.startswith("mrc"):
            return self.forward_mrc(batch, compute_loss)
        if task.startswith("cfp"):
            return self.forward_cfp(batch, compute_loss)
        if task.startswith("og"):
            return self.forward_og(batch, compute_loss)
        raise ValueError("invalid task")

    def forward_mlm(self, batch, compute_loss=True):
        ix = self._index(batch)
        fc = self._fc(self.output_kd and compute_loss)
        o = self.bert(batch, "lang", fc, ix)
        B, Lt, h = o["gmap_embeds"].shape
        head = self.mlm_head.predictions
        rows = ix["mlm_rows"]
        # txt = txt_from_global + txt_from_local, gathered at the masked positions (gather is linear)
        xg = ops.gather_rows(o["gmap_embeds"].reshape(B * Lt, h), rows)
        xl = ops.gather_rows(o["vp_embeds"].reshape(B * Lt, h), rows)
        x = ops.add(xg, xl)
        t = ops.linear(x, head.transform.dense.weight, head.transform.dense.bias, act=L.ACT_GELU)
        t = ops.layer_norm(t, head.transform.LayerNorm.weight, head.transform.LayerNorm.bias,
                           head.transform.LayerNorm.eps)
        logits = ops.linear(t, head.decoder.weight, head.bias, pad_out=True)
        if not compute_loss:
            return {"predict": logits.float() if logits.dtype != torch.float32 else logits}
        loss = ops.cross_entropy(logits, ix["mlm_labels"], -1)
        o.update(loss=loss, logits=logits, predict=logits, row_sample=ix["mlm_row_sample"],
                 loss_inv_n=ix.get("mlm_inv_n"), row_scale=ix.get("mlm_row_scale"),
                 sample_loss=ops.segment_mean(loss.detach(), ix["mlm_row_sample"], ix["mlm_inv_count"], B))
        return o

    def _cls(self, head, x):
        n = head.net
        t = ops.linear(x, n[0].weight, n[0].bias, act=L.ACT_RELU)
        t = ops.layer_norm(t, n[2].weight, n[2].bias, n[2].eps)
        if n[3].weight.shape[0] == 1:
            return ops.rowdot(t, n[3].weight, n[3].bias)
        return ops.linear(t, n[3].weight, n[3].bias)

    def sap_logits(self, o, ix):
        g, v = o["gmap_embeds"], o["vp_embeds"]
        B, G, h = g.shape
        gate = None
        if self.sap_fuse_linear is not None:
            g0 = ops.gather_rows(g.reshape(B * G, h), ix["stop_rows_g"])
            v0 = ops.gather_rows(v.reshape(B * v.shape[1], h), ix["stop_rows_v"])
            gate = self._cls(self.sap_fuse_linear, ops.cat2(g0, v0)).reshape(B)
        g_raw = self._cls(self.global_sap_head, g).reshape(B, G)
        l_raw = self._cls(self.local_sap_head, v).reshape(B, v.shape[1])
        return ops.sap_fuse(g_raw, l_raw, gate, ix)

    def forward_sap(self, batch, compute_loss=True):
        ix = self._index(batch)
        fc = self._fc(self.output_kd and compute_loss)
        o = self.bert(batch, "nav", fc, ix)
        gl, ll, fl = self.sap_logits(o, ix)
        res = dict(global_logits=gl, local_logits=ll, fused_logits=fl, global_act_labels=batch["global_act_labels"],
                   local_act_labels=batch["local_act_labels"])
        if not compute_loss:
            return res
        ga, la = batch["global_act_labels"], batch["local_act_labels"]
        gloss = ops.cross_entropy(gl, ga, -100)
        lloss = ops.cross_entropy(ll, la, -100)
        floss = ops.cross_entropy(fl, ga, -100)
        o.update(res)
        o.update(loss=ops.add(gloss, lloss, floss), sample_loss=floss.detach(), logits=fl)
        return o

    def forward_mrc(self, batch, compute_loss=True):
        """Masked region classification (outputs pinned at train_r2r_magic.py:483-488): the masked views of the
        last-step panorama are zeroed, the local branch's tokens at those views are classified over
        `image_prob_size` classes against the soft labels `vp_view_probs[mask]` with KL."""
        ix = self._index(batch)
        fc = self._fc(False)
        fts = batch.get("traj_view_img_fts")
        if fts is None:
            fts = self.bert.view_features(batch, fc.dtype)
        R, V, Fd = fts.shape
        fts2 = ops.zero_rows_(fts.reshape(R * V, Fd).clone(), ix["mrc_fts_rows"]).view(R, V, Fd)
        o = self.bert(batch, "nav", fc, ix, img_fts=fts2)
        B, Vp, h = o["vp_embeds"].shape
        x = ops.gather_rows(o["vp_embeds"].reshape(B * Vp, h), ix["mrc_rows"])
        logits = self._cls(self.image_classifier, x)
        probs = batch["vp_view_probs"]
        with torch.no_grad():
            targets = ops.gather_rows(probs.reshape(-1, probs.shape[-1]).float(), ix["mrc_tgt_rows"])
        if not compute_loss:
            return (logits.float() if logits.dtype != torch.float32 else logits), targets, None, None
        o.update(loss=ops.soft_cross_entropy(logits, targets), logits=logits, view_targets=targets)
        return o

    def forward_og(self, batch, compute_loss=True):
        """Object grounding (batch schema data/tasks.py:503-559; DUET-lineage head): the object tokens of the last
        panorama, taken from the local branch ([stop] + views + objects), are scored by `og_head`; padded object
        slots get -inf; loss = CE against `obj_labels` (ignore -100, dataset.py:318).
        compute_loss=False returns the `[B, max_objects]` logits."""
        if not hasattr(self, "og_head"):
            raise ValueError("this model was built without the og task (config.pretrain_tasks)")
        ix = self._index(batch)
        if "og_rows" not in ix:
            raise ValueError("og needs a batch with object features (traj_obj_img_fts / traj_vp_obj_lens)")
        fc = self._fc(False)
        o = self.bert(batch, "nav", fc, ix)
        B, Vp, h = o["vp_embeds"].shape
        x = ops.gather_rows(o["vp_embeds"].reshape(B * Vp, h), ix["og_rows"])
        raw = self._cls(self.og_head, x).reshape(B, -1)
        logits = ops.mask_fill(raw, ix["og_valid"])
        if not compute_loss:
            return logits
        o.update(loss=ops.cross_entropy(logits, batch["obj_labels"], -100), logits=logits, obj_logits=logits)
        return o

    def forward_cfp(self, batch, compute_loss=True):
        """Cross-modal feature pooling (outputs pinned at train_r2r_magic.py:545-546): L2-normalised projections of
        the global / local [stop] tokens, their fusion and the text [CLS] token; loss = symmetric InfoNCE of each
        visual feature against the text feature at `cfp_temperature` (validate_cfp, :550-562)."""
        ix = self._index(batch)
        fc = self._fc(False)
        o = self.bert(batch, "nav", fc, ix)
        g3, v3, t3 = o["gmap_embeds"], o["vp_embeds"], o["txt_embeds"]
        B, G, h = g3.shape
        g0 = ops.gather_rows(g3.reshape(B * G, h), ix["stop_rows_g"])
        v0 = ops.gather_rows(v3.reshape(B * v3.shape[1], h), ix["stop_rows_v"])
        t0 = ops.gather_rows(t3.reshape(B * t3.shape[1], h), ix["cls_rows_txt"])
        g = ops.l2norm(ops.linear(g0, self.cfp_gmap_proj.weight, self.cfp_gmap_proj.bias))
        v = ops.l2norm(ops.linear(v0, self.cfp_vp_proj.weight, self.cfp_vp_proj.bias))
        f = ops.l2norm(ops.add(g, v))
        t = ops.l2norm(ops.linear(t0, self.cfp_txt_proj.weight, self.cfp_txt_proj.bias))
        if not compute_loss:
            return tuple(x.float() if x.dtype != torch.float32 else x for x in (g, v, f, t))
        inv_tem = 1.0 / float(self.config.cfp_temperature)
        tgt = ix["arange_b"]

        def nce(a):
            return ops.add(ops.cross_entropy(ops.matmul_nt(a, t, inv_tem), tgt, -100),
                           ops.cross_entropy(ops.matmul_nt(t, a, inv_tem), tgt, -100))

        o.update(loss=ops.add(nce(g), nce(v), nce(f), scale=1.0 / 6.0), cfp_outputs=(g, v, f, t))
        return o
