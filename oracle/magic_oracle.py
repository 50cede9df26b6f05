"""ORACLE (test infrastructure, NOT product code) -- fp32 PyTorch restatement of the MAGIC
pretraining / distillation hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this file.  The product (`vln-magic_b200/`) never does.

PARITY UNPINNED (architecture): the reference omits its model files
(/root/reference/readme.md:75), so the transformer below is OUR frozen
restatement of the public DUET -> GOAT -> MAGIC design.  What IS pinned by
reference code, and followed literally here:
  * module / parameter names          -- pretrain_src/train_r2r_magic.py:189-208 (METER/roberta remap)
  * config attributes                 -- pretrain_src/config/r2r_magic_model_config.json:1-73,
                                         pretrain_src/train_r2r_magic.py:103-160
  * batch schema                      -- pretrain_src/data/tasks.py:110-166 (mlm), :392-451 (sap)
  * forward(batch, task, compute_loss) outputs -- pretrain_src/train_r2r_magic.py:448,483,510-512,545-546
  * KD loss arithmetic                -- pretrain_src/optim/kd_loss.py:5-54 (restated in kd_loss_oracle.py,
                                         pinned against the reference file by tests/golden/kd_loss_*.pt)
  * MAKD aggregation                  -- map_nav_src/r2r/agent.py:546-719 (compute_kd_losses),
                                         :866-869 (MKRW), :1013-1020 (MKTD), :1110-1123 (mix); `makd_losses` is PINNED:
                                         tests/test_makd_agent_pinned.py checks it against tests/golden/makd_agent_ref.pt,
                                         produced by executing the reference function's own source (both roles, all
                                         ability-weight branches, weighted / unweighted, mean / sum)
  * block arithmetic                  -- tests/test_oracle_blocks_pinned.py: BertLayer / BertAttention (self, cross) /
                                         BertEmbeddings vs `transformers`, PanoLayer vs torch.nn.TransformerEncoderLayer
                                         (same state-dict keys, same activations)
Everything else (how the blocks are wired) is tagged [DECISION] where it is made.
"""
import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import kd_loss_oracle as KD

NEG_MASK = -10000.0  # HF "extended attention mask" constant


# ----------------------------------------------------------------------------------------------
# config helpers (train_r2r_magic.py:103-160 restated: prefix stripping teacher_/student_)
# ----------------------------------------------------------------------------------------------
DEFAULT_CONFIG = dict(
    pred_head_dropout_prob=0.1, attention_probs_dropout_prob=0.1, hidden_act="gelu",
    hidden_dropout_prob=0.1, hidden_size=768, initializer_range=0.02, intermediate_size=3072,
    num_l_layers=6, num_x_layers=3, num_pano_layers=2, layer_norm_eps=1e-12,
    max_position_embeddings=514, max_action_steps=100, num_attention_heads=12, type_vocab_size=1,
    update_lang_bert=True, vocab_size=50265, use_lang2visn_attn=True, graph_sprels=True,
    glocal_fuse=True, image_feat_size=768, image_prob_size=1000, angle_feat_size=4, obj_feat_size=0,
    adaptive_pano_fusion=True, cfp_temperature=1.0,
)


def make_config(hidden_size, num_l_layers=6, num_x_layers=3, num_pano_layers=2, mlp_ratio=4,
                role="student", teacher_hidden_size=None, pretrain_tasks=("mlm", "sap"), **over):
    """Equivalent of the student/teacher config derivation at train_r2r_magic.py:125-160:
    intermediate = hidden * mlp_ratio, heads = hidden / 64."""
    cfg = dict(DEFAULT_CONFIG)
    cfg.update(hidden_size=hidden_size, num_l_layers=num_l_layers, num_x_layers=num_x_layers,
               num_pano_layers=num_pano_layers, intermediate_size=int(hidden_size * mlp_ratio),
               num_attention_heads=int(hidden_size / 64), role=role, kd=teacher_hidden_size is not None,
               pretrain_tasks=set(pretrain_tasks))
    if teacher_hidden_size is not None:
        cfg["teacher_hidden_size"] = teacher_hidden_size
    cfg.update(over)
    return SimpleNamespace(**cfg)


def gen_seq_masks(lens, max_len=None):
    """True = valid.  Semantics of pretrain_src/data/common.py:60-75."""
    if max_len is None:
        max_len = int(lens.max())
    return torch.arange(max_len, device=lens.device)[None, :] < lens[:, None]


def ext_mask(mask):
    """[B,N] bool -> additive [B,1,1,N]."""
    return (~mask).to(torch.float32)[:, None, None, :] * NEG_MASK


# ----------------------------------------------------------------------------------------------
# BERT blocks (HF BertLayer naming; train_r2r_magic.py:190-202 copies keys verbatim)
# ----------------------------------------------------------------------------------------------
class BertEmbeddings(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.word_embeddings = nn.Embedding(c.vocab_size, c.hidden_size)
        self.position_embeddings = nn.Embedding(c.max_position_embeddings, c.hidden_size)
        self.token_type_embeddings = nn.Embedding(c.type_vocab_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.dropout = nn.Dropout(c.hidden_dropout_prob)

    def forward(self, ids):
        B, L = ids.shape
        pos = torch.arange(L, device=ids.device)  # [DECISION] position ids = arange(L) (A.2)
        e = self.word_embeddings(ids) + self.position_embeddings(pos)[None] + self.token_type_embeddings.weight[0]
        return self.dropout(self.LayerNorm(e))


class BertSelfAttention(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.H = c.num_attention_heads
        self.d = c.hidden_size // self.H
        self.query = nn.Linear(c.hidden_size, c.hidden_size)
        self.key = nn.Linear(c.hidden_size, c.hidden_size)
        self.value = nn.Linear(c.hidden_size, c.hidden_size)
        self.dropout = nn.Dropout(c.attention_probs_dropout_prob)

    def forward(self, x, ctx, add_mask):
        B, Lq, _ = x.shape
        Lk = ctx.shape[1]
        q = self.query(x).view(B, Lq, self.H, self.d).transpose(1, 2)
        k = self.key(ctx).view(B, Lk, self.H, self.d).transpose(1, 2)
        v = self.value(ctx).view(B, Lk, self.H, self.d).transpose(1, 2)
        s = q @ k.transpose(-1, -2) / math.sqrt(self.d)
        if add_mask is not None:
            s = s + add_mask
        p = torch.softmax(s, -1)
        o = (self.dropout(p) @ v).transpose(1, 2).reshape(B, Lq, self.H * self.d)
        return o, p.mean(1)  # [DECISION] KD attention map = head-mean of pre-dropout probs (A.1)


class BertSelfOutput(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.dropout = nn.Dropout(c.hidden_dropout_prob)

    def forward(self, h, res):
        return self.LayerNorm(self.dropout(self.dense(h)) + res)


class BertAttention(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.self = BertSelfAttention(c)
        self.output = BertSelfOutput(c)

    def forward(self, x, ctx, add_mask):
        o, p = self.self(x, ctx, add_mask)
        return self.output(o, x), p


class BertIntermediate(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.intermediate_size)

    def forward(self, x):
        return F.gelu(self.dense(x))


class BertOutput(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.intermediate_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.dropout = nn.Dropout(c.hidden_dropout_prob)

    def forward(self, h, res):
        return self.LayerNorm(self.dropout(self.dense(h)) + res)


class BertLayer(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.attention = BertAttention(c)
        self.intermediate = BertIntermediate(c)
        self.output = BertOutput(c)

    def forward(self, x, add_mask):
        a, p = self.attention(x, x, add_mask)
        return self.output(self.intermediate(a), a), p


class LangEncoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.layer = nn.ModuleList([BertLayer(c) for _ in range(c.num_l_layers)])


class BertCrossLayer(nn.Module):
    """METER BertCrossLayer (key names pinned train_r2r_magic.py:203-206).
    [DECISION] order self-attn -> cross-attn -> FFN, each post-LN."""

    def __init__(self, c):
        super().__init__()
        self.attention = BertAttention(c)
        self.crossattention = BertAttention(c)
        self.intermediate = BertIntermediate(c)
        self.output = BertOutput(c)

    def forward(self, x, ctx, x_add_mask, ctx_add_mask):
        a, p_self = self.attention(x, x, x_add_mask)
        cx, p_cross = self.crossattention(a, ctx, ctx_add_mask)
        out = self.output(self.intermediate(cx), cx)
        return out, torch.cat([p_self, p_cross], -1)  # [DECISION] attn map = [self | cross] on last dim


class CrossEncoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.crossattention = nn.ModuleList([BertCrossLayer(c) for _ in range(c.num_x_layers)])

    def forward(self, x, ctx, x_mask, ctx_mask, sprel_bias=None):
        xm = ext_mask(x_mask)
        if sprel_bias is not None:
            xm = xm + sprel_bias
        cm = ext_mask(ctx_mask)
        attns = []
        for layer in self.crossattention:
            x, p = layer(x, ctx, xm, cm)
            attns.append(p)
        return x, torch.stack(attns, 1)


# ----------------------------------------------------------------------------------------------
# panorama encoder: pre-LN layers, DETR/torch.nn.TransformerEncoderLayer naming
# ----------------------------------------------------------------------------------------------
class MHA(nn.Module):
    def __init__(self, h, H):
        super().__init__()
        self.H, self.d = H, h // H
        self.in_proj_weight = nn.Parameter(torch.empty(3 * h, h))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * h))
        self.out_proj = nn.Linear(h, h)

    def forward(self, x, key_mask, drop):
        B, N, h = x.shape
        qkv = F.linear(x, self.in_proj_weight, self.in_proj_bias).view(B, N, 3, self.H, self.d)
        q, k, v = qkv[:, :, 0].transpose(1, 2), qkv[:, :, 1].transpose(1, 2), qkv[:, :, 2].transpose(1, 2)
        s = q @ k.transpose(-1, -2) / math.sqrt(self.d)
        s = s.masked_fill(~key_mask[:, None, None, :], float("-inf"))  # key_padding_mask semantics
        p = torch.softmax(s, -1)
        o = (drop(p) @ v).transpose(1, 2).reshape(B, N, h)
        return self.out_proj(o), p.mean(1)


class PanoLayer(nn.Module):
    def __init__(self, c):
        super().__init__()
        h = c.hidden_size
        self.self_attn = MHA(h, c.num_attention_heads)
        self.linear1 = nn.Linear(h, c.intermediate_size)
        self.linear2 = nn.Linear(c.intermediate_size, h)
        self.norm1 = nn.LayerNorm(h, eps=c.layer_norm_eps)
        self.norm2 = nn.LayerNorm(h, eps=c.layer_norm_eps)
        self.dropout = nn.Dropout(c.hidden_dropout_prob)
        self.dropout1 = nn.Dropout(c.hidden_dropout_prob)
        self.dropout2 = nn.Dropout(c.hidden_dropout_prob)

    def forward(self, x, key_mask):
        a, p = self.self_attn(self.norm1(x), key_mask, self.dropout)
        x = x + self.dropout1(a)
        f = self.linear2(self.dropout(F.gelu(self.linear1(self.norm2(x)))))
        return x + self.dropout2(f), p


class PanoEncoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.layers = nn.ModuleList([PanoLayer(c) for _ in range(c.num_pano_layers)])
        self.norm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)


class ImageEmbeddings(nn.Module):
    def __init__(self, c):
        super().__init__()
        h = c.hidden_size
        self.img_linear = nn.Linear(c.image_feat_size, h)
        self.img_layer_norm = nn.LayerNorm(h, eps=c.layer_norm_eps)
        self.loc_linear = nn.Linear(c.angle_feat_size + 3, h)
        self.loc_layer_norm = nn.LayerNorm(h, eps=c.layer_norm_eps)
        self.nav_type_embedding = nn.Embedding(3, h)
        if getattr(c, "obj_feat_size", 0) > 0:  # object tokens (REVERIE / SOON), DUET-lineage names
            self.obj_linear = nn.Linear(c.obj_feat_size, h)
            self.obj_layer_norm = nn.LayerNorm(h, eps=c.layer_norm_eps)
        else:
            self.obj_linear = self.obj_layer_norm = None
        self.layer_norm = nn.LayerNorm(h, eps=c.layer_norm_eps)
        self.dropout = nn.Dropout(c.hidden_dropout_prob)
        self.pano_encoder = PanoEncoder(c) if c.num_pano_layers > 0 else None
        # [DECISION] adaptive pano fusion = learned attention pooling over the valid views (A.2)
        self.adaptive_pano_attn = nn.Linear(h, 1) if c.adaptive_pano_fusion else None

    def forward(self, img_fts, loc_fts, nav_types, view_lens, type_embed, obj_fts=None, obj_lens=None):
        img = self.img_layer_norm(self.img_linear(img_fts))
        if obj_fts is not None and self.obj_linear is not None:
            # tokens of a panorama = its views followed by its objects (dataset.py:447,494-508), padded to the longest
            obj = self.obj_layer_norm(self.obj_linear(obj_fts))
            rows = [torch.cat([img[r, :int(view_lens[r])], obj[r, :int(obj_lens[r])]], 0) for r in range(img.shape[0])]
            n = loc_fts.shape[1]
            img = torch.stack([F.pad(x, (0, 0, 0, n - x.shape[0])) for x in rows], 0)
            view_lens = view_lens + obj_lens
        e = img + self.loc_layer_norm(self.loc_linear(loc_fts)) + self.nav_type_embedding(nav_types) + type_embed
        e = self.dropout(self.layer_norm(e))
        mask = gen_seq_masks(view_lens, e.shape[1])
        attns = []
        if self.pano_encoder is not None:
            for layer in self.pano_encoder.layers:
                e, p = layer(e, mask)
                attns.append(p)
            e = self.pano_encoder.norm(e)
        if self.adaptive_pano_attn is not None:
            s = self.adaptive_pano_attn(e).squeeze(-1).masked_fill(~mask, float("-inf"))
            fused = (torch.softmax(s, -1)[..., None] * e).sum(1)
        else:
            m = mask[..., None].to(e.dtype)
            fused = (e * m).sum(1) / view_lens[:, None].to(e.dtype)
        attns = torch.stack(attns, 1) if attns else e.new_zeros(e.shape[0], 0, e.shape[1], e.shape[1])
        return e, fused, attns


class LocalVPEncoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.vp_pos_embeddings = nn.Sequential(
            nn.Linear(c.angle_feat_size * 2 + 6, c.hidden_size), nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps))
        self.encoder = CrossEncoder(c)


class GlobalMapEncoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.gmap_pos_embeddings = nn.Sequential(
            nn.Linear(c.angle_feat_size + 3, c.hidden_size), nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps))
        self.gmap_step_embeddings = nn.Embedding(c.max_action_steps, c.hidden_size)
        self.encoder = CrossEncoder(c)
        self.sprel_linear = nn.Linear(1, 1) if c.graph_sprels else None


# ----------------------------------------------------------------------------------------------
# graph index logic on vp-id STRINGS -- python loops, the spec the device path
# (integer index tensors) must reproduce bit-exactly
# ----------------------------------------------------------------------------------------------
def aggregate_gmap_features(pano_embeds, pano_fused, batch):
    """Returns [B, G, h] node features; node 0 = [stop] = zeros (A.4)."""
    B = len(batch["traj_step_lens"])
    G = batch["gmap_step_ids"].shape[1]
    h = pano_embeds.shape[-1]
    out = []
    row0 = 0
    for b in range(B):
        T = batch["traj_step_lens"][b]
        visited = {}
        unvisited = {}
        for t in range(T):
            visited[batch["traj_vpids"][b][t]] = pano_fused[row0 + t]  # last occurrence wins
            for j, c in enumerate(batch["traj_cand_vpids"][b][t]):
                unvisited.setdefault(c, []).append(pano_embeds[row0 + t, j])
        feats = [pano_embeds.new_zeros(h)]
        for v in batch["gmap_vpids"][b][1:]:
            if v in visited:
                feats.append(visited[v])
            else:
                feats.append(torch.stack(unvisited[v], 0).mean(0))
        feats = torch.stack(feats, 0)
        out.append(F.pad(feats, (0, 0, 0, G - feats.shape[0])))
        row0 += T
    return torch.stack(out, 0)


def vp_lens_of(batch):
    """Number of valid local tokens = tokens (views + objects) of the LAST step + 1 ([stop]) -- DUET lineage computes this inside
    the model from traj_vp_view_lens.  NB the reference collate's own batch['vp_lens'] is `len(x[-1])` of a
    [Vp,14] tensor, i.e. the constant 14 (pretrain_src/data/tasks.py:153); it is not usable as a length."""
    rows = last_step_rows(batch)
    lens = batch["traj_vp_view_lens"][rows]
    if batch.get("traj_vp_obj_lens") is not None:
        lens = lens + batch["traj_vp_obj_lens"][rows]
    return lens + 1


def last_step_rows(batch):
    rows, r = [], 0
    for T in batch["traj_step_lens"]:
        r += T
        rows.append(r - 1)
    return rows


def fuse_sap_logits(global_logits, local_logits, batch):
    """SURVEY.md A.4 'SAP fusion' (DUET lineage)."""
    fused = global_logits.clone()
    fused[:, 0] = fused[:, 0] + local_logits[:, 0]
    for b in range(global_logits.shape[0]):
        vpids = batch["gmap_vpids"][b]
        vis = batch["gmap_visited_masks"][b]
        visited_nodes = set(vp for vp, m in zip(vpids, vis.tolist()) if m)
        tmp, bw = {}, 0
        for j, c in enumerate(batch["traj_cand_vpids"][b][-1]):
            if c in visited_nodes:
                bw = bw + local_logits[b, j + 1]
            else:
                tmp[c] = local_logits[b, j + 1]
        for n, vp in enumerate(vpids):
            if n > 0 and vp not in visited_nodes:
                fused[b, n] = fused[b, n] + (tmp[vp] if vp in tmp else bw)
    return fused


# ----------------------------------------------------------------------------------------------
# backbone
# ----------------------------------------------------------------------------------------------
class GlocalTextPathCMT(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.config = c
        self.embeddings = BertEmbeddings(c)
        self.lang_encoder = LangEncoder(c)
        self.img_embeddings = ImageEmbeddings(c)
        self.local_encoder = LocalVPEncoder(c)
        self.global_encoder = GlobalMapEncoder(c)
        ht = getattr(c, "teacher_hidden_size", None)
        if getattr(c, "role", "student") == "student" and getattr(c, "kd", False) and ht:
            # student-side KD projection heads (agent.py:568,600-601,643,661; agent_base.py:330)
            for name in ("txt_emb_w", "kdl_img_w", "kdl_avg_img_w", "global_cross_w", "local_cross_w",
                         "vp_txt_w", "gmap_txt_w"):
                setattr(self, name, nn.Linear(c.hidden_size, ht))
            kdl = getattr(c, "kdl", None) or {}
            kind = kdl.get("kdl_adaptive_ability_weight_type") if hasattr(kdl, "get") else \
                getattr(kdl, "kdl_adaptive_ability_weight_type", None)
            if kind == "learned_weight":
                # learned ability weights used through softplus (agent.py:585,618,678,681,713)
                # [DECISION] scalars initialised so that softplus(.) = 1
                for name in ("kdl_txt_weight", "kdl_img_weight", "kdl_global_weight", "kdl_local_weight",
                             "kdl_predict_weight"):
                    setattr(self, name, nn.Parameter(torch.full((1,), 0.5413248546129181)))

    def forward_text(self, txt_ids, txt_lens):
        x = self.embeddings(txt_ids)
        mask = gen_seq_masks(txt_lens, txt_ids.shape[1])
        am = ext_mask(mask)
        attns = []
        for layer in self.lang_encoder.layer:
            x, p = layer(x, am)
            attns.append(p)
        if not self.config.update_lang_bert:
            x = x.detach()
        return x, mask, torch.stack(attns, 1)

    def forward_pano(self, batch):
        type_embed = self.embeddings.token_type_embeddings.weight[0]  # [DECISION] type index 0 (A.2)
        return self.img_embeddings(batch["traj_view_img_fts"], batch["traj_loc_fts"], batch["traj_nav_types"],
                                   batch["traj_vp_view_lens"], type_embed, batch.get("traj_obj_img_fts"),
                                   batch.get("traj_vp_obj_lens"))

    def gmap_input(self, pano_embeds, pano_fused, batch):
        ge = self.global_encoder
        x = aggregate_gmap_features(pano_embeds, pano_fused, batch)
        x = x + ge.gmap_step_embeddings(batch["gmap_step_ids"]) + ge.gmap_pos_embeddings(batch["gmap_pos_fts"])
        mask = gen_seq_masks(batch["gmap_lens"], x.shape[1])
        sprel = None
        if ge.sprel_linear is not None:
            sprel = ge.sprel_linear(batch["gmap_pair_dists"].unsqueeze(3)).squeeze(3).unsqueeze(1)
        return x, mask, sprel

    def vp_input(self, pano_embeds, batch):
        rows = last_step_rows(batch)
        last = pano_embeds[rows]
        Vp = batch["vp_pos_fts"].shape[1]
        x = torch.cat([torch.zeros_like(last[:, :1]), last], 1)[:, :Vp]
        x = x + self.local_encoder.vp_pos_embeddings(batch["vp_pos_fts"])
        mask = gen_seq_masks(vp_lens_of(batch), Vp)
        return x, mask

    def forward(self, batch, mode):
        """mode 'nav': visual queries attend text (SAP/CFP/MRC);  mode 'lang': text queries attend the
        gmap / vp INPUT embeddings through the same x-layers and the two results are summed
        ([DECISION] use_lang2visn_attn branch)."""
        out = {}
        txt, txt_mask, txt_attns = self.forward_text(batch["txt_ids"], batch["txt_lens"])
        pano, pano_fused, img_attns = self.forward_pano(batch)
        out.update(txt_embeds=txt, txt_attns=txt_attns, pano_embeds=pano, pano_fused_embeds=pano_fused,
                   img_attns=img_attns, txt_masks=txt_mask)
        g_in, g_mask, sprel = self.gmap_input(pano, pano_fused, batch)
        v_in, v_mask = self.vp_input(pano, batch)
        if mode == "nav":
            g, g_attn = self.global_encoder.encoder(g_in, txt, g_mask, txt_mask, sprel)
            v, v_attn = self.local_encoder.encoder(v_in, txt, v_mask, txt_mask)
        else:
            g, g_attn = self.global_encoder.encoder(txt, g_in, txt_mask, g_mask)
            v, v_attn = self.local_encoder.encoder(txt, v_in, txt_mask, v_mask)
        out.update(gmap_embeds=g, gmap_attns=g_attn, vp_embeds=v, vp_attns=v_attn)
        return out


# ----------------------------------------------------------------------------------------------
# heads
# ----------------------------------------------------------------------------------------------
class BertPredictionHeadTransform(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)

    def forward(self, x):
        return self.LayerNorm(F.gelu(self.dense(x)))


class BertLMPredictionHead(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.transform = BertPredictionHeadTransform(c)
        self.decoder = nn.Linear(c.hidden_size, c.vocab_size, bias=False)
        self.bias = nn.Parameter(torch.zeros(c.vocab_size))

    def forward(self, x):
        return self.decoder(self.transform(x)) + self.bias


class BertOnlyMLMHead(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.predictions = BertLMPredictionHead(c)

    def forward(self, x):
        return self.predictions(x)


class ClsPrediction(nn.Module):
    def __init__(self, h, input_size=None, out=1, eps=1e-12):
        super().__init__()
        input_size = h if input_size is None else input_size
        self.net = nn.Sequential(nn.Linear(input_size, h), nn.ReLU(), nn.LayerNorm(h, eps=eps), nn.Linear(h, out))

    def forward(self, x):
        return self.net(x)


class GlocalTextPathCMTPreTraining(nn.Module):
    def __init__(self, config):
        super().__init__()
        c = self.config = config
        self.bert = GlocalTextPathCMT(c)
        tasks = c.pretrain_tasks
        if "mlm" in tasks:
            self.mlm_head = BertOnlyMLMHead(c)
            self.mlm_head.predictions.decoder.weight = self.bert.embeddings.word_embeddings.weight  # tied
        if "sap" in tasks or "cfp" in tasks:
            self.global_sap_head = ClsPrediction(c.hidden_size, eps=c.layer_norm_eps)
            self.local_sap_head = ClsPrediction(c.hidden_size, eps=c.layer_norm_eps)
            self.sap_fuse_linear = ClsPrediction(c.hidden_size, input_size=2 * c.hidden_size,
                                                 eps=c.layer_norm_eps) if c.glocal_fuse else None
        if "mrc" in tasks:
            self.image_classifier = ClsPrediction(c.hidden_size, out=c.image_prob_size, eps=c.layer_norm_eps)
        if "cfp" in tasks:
            for n in ("cfp_gmap_proj", "cfp_vp_proj", "cfp_txt_proj"):
                setattr(self, n, nn.Linear(c.hidden_size, c.hidden_size))
        if "og" in tasks:
            self.og_head = ClsPrediction(c.hidden_size, eps=c.layer_norm_eps)
        self.apply(self._init)

    def _init(self, m):
        r = self.config.initializer_range
        if isinstance(m, (nn.Linear, nn.Embedding)):
            m.weight.data.normal_(0.0, r)
        elif isinstance(m, nn.LayerNorm):
            m.weight.data.fill_(1.0)
            m.bias.data.zero_()
        elif isinstance(m, MHA):
            m.in_proj_weight.data.normal_(0.0, r)
            m.in_proj_bias.data.zero_()
        if isinstance(m, nn.Linear) and m.bias is not None:
            m.bias.data.zero_()

    # -- tasks --------------------------------------------------------------------------------
    def forward(self, batch, task, compute_loss=True):
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

    def forward_og(self, batch, compute_loss):
        """[INFERRED, DUET lineage forward_og] object tokens of the last panorama, read from the local branch
        (vp_embeds[b, 1 + view_len : 1 + view_len + obj_len]), scored by og_head, -inf on padded slots, CE with
        ignore_index -100 (dataset.py:318)."""
        o = self.bert(batch, "nav")
        rows = last_step_rows(batch)
        vl, ol = batch["traj_vp_view_lens"][rows], batch["traj_vp_obj_lens"][rows]
        O = batch["traj_obj_img_fts"].shape[1]
        v = o["vp_embeds"]
        obj = torch.stack([F.pad(v[b, 1 + int(vl[b]): 1 + int(vl[b]) + int(ol[b])], (0, 0, 0, O - int(ol[b])))
                           for b in range(v.shape[0])], 0)
        logits = self.og_head(obj).squeeze(2).masked_fill(~gen_seq_masks(ol, O), float("-inf"))
        if not compute_loss:
            return logits
        o.update(loss=F.cross_entropy(logits, batch["obj_labels"], reduction="none", ignore_index=-100), logits=logits)
        return o

    def forward_mlm(self, batch, compute_loss):
        o = self.bert(batch, "lang")
        txt = o["gmap_embeds"] + o["vp_embeds"]
        sel = batch["txt_labels"] != -1  # row-major nonzero order (train_r2r_magic.py:450-452)
        logits = self.mlm_head(txt[sel])
        if not compute_loss:
            return {"predict": logits}
        labels = batch["txt_labels"][sel]
        loss = F.cross_entropy(logits, labels, reduction="none")
        rows = sel.nonzero()[:, 0]
        B = sel.shape[0]
        cnt = torch.zeros(B, device=loss.device).index_add_(0, rows, torch.ones_like(loss))
        sample_loss = torch.zeros(B, device=loss.device).index_add_(0, rows, loss) / cnt.clamp(min=1)
        o.update(loss=loss, sample_loss=sample_loss, logits=logits, predict=logits, row_sample=rows)
        return o

    def sap_logits(self, o, batch):
        g, v = o["gmap_embeds"], o["vp_embeds"]
        if self.sap_fuse_linear is None:
            fw = 0.5
        else:
            fw = torch.sigmoid(self.sap_fuse_linear(torch.cat([g[:, 0], v[:, 0]], 1)))
        gl = self.global_sap_head(g).squeeze(2) * fw
        gmask = gen_seq_masks(batch["gmap_lens"], g.shape[1])
        gl = gl.masked_fill(batch["gmap_visited_masks"], float("-inf")).masked_fill(~gmask, float("-inf"))
        ll = self.local_sap_head(v).squeeze(2) * (1 - fw)
        rows = last_step_rows(batch)
        Vp = v.shape[1]
        nav = torch.cat([torch.ones_like(batch["traj_nav_types"][rows][:, :1], dtype=torch.bool),
                         batch["traj_nav_types"][rows] == 1], 1)[:, :Vp]
        nav = nav & gen_seq_masks(vp_lens_of(batch), Vp)
        ll = ll.masked_fill(~nav, float("-inf"))
        fl = fuse_sap_logits(gl, ll, batch)
        return gl, ll, fl

    def forward_sap(self, batch, compute_loss):
        o = self.bert(batch, "nav")
        gl, ll, fl = self.sap_logits(o, batch)
        res = dict(global_logits=gl, local_logits=ll, fused_logits=fl,
                   global_act_labels=batch["global_act_labels"], local_act_labels=batch["local_act_labels"])
        if not compute_loss:
            return res
        ga, la = batch["global_act_labels"], batch["local_act_labels"]
        gloss = F.cross_entropy(gl, ga, reduction="none", ignore_index=-100)
        lloss = F.cross_entropy(ll, la, reduction="none", ignore_index=-100)
        floss = F.cross_entropy(fl, ga, reduction="none", ignore_index=-100)
        o.update(res)
        o.update(loss=gloss + lloss + floss, sample_loss=floss, logits=fl)
        return o

    def forward_mrc(self, batch, compute_loss):
        # mask the last-step views, classify them from the local branch (DUET lineage)
        rows = last_step_rows(batch)
        m = batch["vp_view_mrc_masks"]
        fts = batch["traj_view_img_fts"].clone()
        fts[rows] = fts[rows].masked_fill(m[..., None], 0)
        b2 = dict(batch)
        b2["traj_view_img_fts"] = fts
        o = self.bert(b2, "nav")
        v = o["vp_embeds"][:, 1:]
        logits = self.image_classifier(v[m[:, :v.shape[1]]])
        targets = batch["vp_view_probs"][m]
        if not compute_loss:
            return logits, targets, None, None
        loss = F.kl_div(F.log_softmax(logits, -1), targets, reduction="none").sum(1)
        o.update(loss=loss)
        return o

    def forward_cfp(self, batch, compute_loss):
        o = self.bert(batch, "nav")
        g = F.normalize(self.cfp_gmap_proj(o["gmap_embeds"][:, 0]), dim=-1)
        v = F.normalize(self.cfp_vp_proj(o["vp_embeds"][:, 0]), dim=-1)
        f = F.normalize(g + v, dim=-1)
        t = F.normalize(self.cfp_txt_proj(o["txt_embeds"][:, 0]), dim=-1)
        if not compute_loss:
            return g, v, f, t
        tem = self.config.cfp_temperature
        tgt = torch.arange(g.shape[0], device=g.device)

        def nce(a):
            s = a @ t.T / tem
            return (F.cross_entropy(s, tgt, reduction="none") + F.cross_entropy(s.T, tgt, reduction="none")) / 2

        o.update(loss=(nce(g) + nce(v) + nce(f)) / 3.0)
        return o


# ----------------------------------------------------------------------------------------------
# MAKD (pretraining composition of the pinned pieces)
# ----------------------------------------------------------------------------------------------
KDL_DEFAULT = dict(kd_alpha=0.5, kd_temperature=2.0, rw_temp=4.0, t_sample_preprocess="exp",
                   t_sample_preprocess_exp_decay=0.7, teacher_sample_hard_mining=True,
                   kdl_adaptive_ability_weight=True, kdl_adaptive_ability_weight_type="RW",
                   kdl_tasks=("txt", "img", "local", "global", "predict"), kdl_task_types=("emb", "attn"))


def mkrw_weights(gen=None, rw_temp=4.0, device="cpu"):
    """agent.py:866-869: softmax(randn(5)/rw_temp)*5, order [txt, img, global, local, predict]."""
    return torch.softmax(torch.randn(5, generator=gen).to(device) / rw_temp, 0) * 5


def mktd_weights(t_sample_loss, decay=0.7, preprocess="exp"):
    """agent.py:1013-1020 with optim/kd_loss.py:43-44 ('exp') or :46-54 ('norm'; agent_base.py:172-175)."""
    if preprocess == "norm":
        return KD.invert_normalized_losses(t_sample_loss.detach(), decay_rate=decay)
    return KD.exponential_decay(t_sample_loss.detach(), decay_rate=decay)


def ability_weights(student, k, rw, weight_owner=None):
    """-> (5 multipliers, divisor of the two image embedding losses): agent.py:583-593, 616-625, 675-693, 710-717.
    `weight_owner`: the model whose learned kdl_*_weight parameters scale the losses -- `s_model` of agent.py:553,557,
    i.e. the LEARNER: the small model for role t2s (default), the large model for role s2t."""
    if not k.get("kdl_adaptive_ability_weight", True):
        return [1.0] * 5, 2.0
    kind = k.get("kdl_adaptive_ability_weight_type", "RW")
    if kind == "learned_weight":
        b = (weight_owner if weight_owner is not None else student).bert
        return [F.softplus(getattr(b, n)).squeeze(0) for n in ("kdl_txt_weight", "kdl_img_weight", "kdl_global_weight",
                                                               "kdl_local_weight", "kdl_predict_weight")], 2.0
    return rw, 1.0


def makd_losses(student, s_out, t_out, task, rw, t_w, kdl=None, role="t2s", weight_owner=None):
    """agent.py:546-719 with the pretrain reductions of optim/kd_loss.py.  `student` is always the SMALL model: it
    owns the up-projections (txt_emb_w, kdl_img_w, kdl_avg_img_w, global_cross_w, local_cross_w).
      role 't2s' (agent.py:550-552): learner = small model; prediction = proj(s_out), target = t_out.detach().
      role 's2t' (agent.py:553-556, ICoD): learner = LARGE model; the caller passes (s_out, t_out) =
        (large model's outputs, small model's outputs) exactly like agent.py:1022; prediction = s_out as is,
        target = proj(t_out).detach() (agent.py:571,605-606,647,665); `t_w` are then the SMALL model's MKTD
        weights (agent.py:1009-1011 stored under the same 'sample_weights' key).
    Reductions: `kd_loss_type` None (default) = the pretraining file's `mean` semantics (silent fallback on a weight /
    batch mismatch); 'mean' / 'sum' = the fine-tune file's (map_nav_src/utils/kd_loss.py; agent.py:554 for role t2s,
    always 'mean' for role s2t, :557).  PINNED: tests/test_makd_agent_pinned.py runs this function against
    tests/golden/makd_agent_ref.pt, which the reference's own compute_kd_losses SOURCE produced.
    Returns dict of the 10 named scalars (agent.py:824-835 names)."""
    k = dict(KDL_DEFAULT)
    if kdl:
        k.update(kdl)
    bert = student.bert
    rw, img_div = ability_weights(student, k, rw, weight_owner)
    lt = k.get("kd_loss_type")
    if role == "s2t" and lt is not None:
        lt = "mean"
    T = k["kd_temperature"]
    emb = "emb" in k["kdl_task_types"]
    att = "attn" in k["kdl_task_types"]
    L = {}
    z = s_out["txt_embeds"].new_zeros(())
    min_len = min(s_out["txt_attns"].shape[1], t_out["txt_attns"].shape[1])  # agent.py:560
    # [DECISION] x-layer maps are additionally clipped to their own common depth (the reference expression
    # `[:, :min_len]` would raise on teacher/student pairs with different num_x_layers, e.g. 4 vs 3)
    nx = min(min_len, s_out["gmap_attns"].shape[1], t_out["gmap_attns"].shape[1])

    def pair(proj, key):
        if role == "t2s":
            return proj(s_out[key]), t_out[key].detach()
        return s_out[key], proj(t_out[key]).detach()

    def e(proj, key, ri):
        return KD.mse_loss(*pair(proj, key), t_w, loss_type=lt) * rw[ri] if emb else z

    def a(key, n, ri):
        sa, ta = (s_out[key], t_out[key]) if n is None else (s_out[key][:, :n], t_out[key][:, :n])
        return KD.mse_loss(sa, ta.detach(), t_w, loss_type=lt) * rw[ri] if att else z

    if "txt" in k["kdl_tasks"]:
        L["txt_emb_loss"] = e(bert.txt_emb_w, "txt_embeds", 0)
        L["txt_attn_loss"] = a("txt_attns", min_len, 0)
    if "img" in k["kdl_tasks"]:
        # agent.py:620-622: under RW the two image emb losses are NOT halved; :618-619, :624-625 otherwise halved
        L["img_emb_loss"] = e(bert.kdl_img_w, "pano_embeds", 1) / img_div
        L["avg_img_emb_loss"] = e(bert.kdl_avg_img_w, "pano_fused_embeds", 1) / img_div
        L["img_attn_loss"] = a("img_attns", None, 1)  # agent.py:628 unsliced
    gw, lw = (bert.gmap_txt_w, bert.vp_txt_w) if task.startswith("mlm") else (bert.global_cross_w, bert.local_cross_w)
    if "global" in k["kdl_tasks"]:
        L["global_emb_loss"] = e(gw, "gmap_embeds", 2)
        L["global_attn_loss"] = a("gmap_attns", nx, 2)
    if "local" in k["kdl_tasks"]:
        L["local_emb_loss"] = e(lw, "vp_embeds", 3)
        L["local_attn_loss"] = a("vp_attns", nx, 3)
    if "predict" in k["kdl_tasks"]:
        w = t_w
        if w is not None and task.startswith("mlm"):
            w = t_w[s_out["row_sample"]]  # [DECISION] per-row weights for the [n_masked, vocab] logits
        L["predict_loss"] = KD.kd_loss(s_out["logits"], t_out["logits"].detach(), temperature=T, t_sample_weights=w,
                                       loss_type=lt) * rw[4]
    return L


def distill_step_loss(student, teacher, batch, task, rw, kdl=None):
    """The (missing) step, SURVEY.md 3.2: returns (total, sup, kd_total, dict)."""
    k = dict(KDL_DEFAULT)
    if kdl:
        k.update(kdl)
    with torch.no_grad():
        t_out = teacher(batch, task, True)
    s_out = student(batch, task, True)
    sup = s_out["loss"].mean()
    t_w = mktd_weights(t_out["sample_loss"], k["t_sample_preprocess_exp_decay"], k["t_sample_preprocess"]) \
        if k["teacher_sample_hard_mining"] else None
    L = makd_losses(student, s_out, t_out, task, rw, t_w, k)
    kd_total = sum(L.values())
    total = k["kd_alpha"] * kd_total + (1 - k["kd_alpha"]) * sup  # agent.py:1119
    return total, sup, kd_total, L, s_out, t_out


def icod_step_loss(student, teacher, batch, task, rw, t_rw, kdl=None):
    """ICoD co-update (`--train_kdl_teacher`, agent.py:1019-1022, 1136-1149; agent_base.py:260-279): one step
    yields TWO losses.  The small model learns from the large one exactly as in `distill_step_loss` (role t2s); the
    large model, whose forward now runs WITH grad, learns from the small one (role s2t: targets are the small
    model's outputs pushed through the small model's projections, detached) mixed with its own supervised loss
    by t_kdl_alpha (parser.py:184: 0.5).  Under RW the large model's ability weights are the small model's draw
    of the step (agent.py:869-871: t_softmax_weights = s_softmax_weights); `t_rw` is kept separate for the 'grad' mode.
    [DECISION] the s2t KD sum enters like the t2s one (mean-reduced terms, no extra batch factor): the fine-tune
    rollout's `* train_ml` (agent.py:1143) is a sum-over-steps convention that pretraining's mean losses lack.
    Returns (total_s, total_t, dict_s, dict_t, s_out, t_out)."""
    k = dict(KDL_DEFAULT)
    k.setdefault("t_kd_alpha", 0.5)
    if kdl:
        k.update(kdl)
    t_out = teacher(batch, task, True)
    s_out = student(batch, task, True)
    pre = k["t_sample_preprocess"]
    t_w = mktd_weights(t_out["sample_loss"], k["t_sample_preprocess_exp_decay"], pre) if k["teacher_sample_hard_mining"] else None
    s_w = mktd_weights(s_out["sample_loss"], k["t_sample_preprocess_exp_decay"], pre) if k["teacher_sample_hard_mining"] else None
    Ls = makd_losses(student, s_out, t_out, task, rw, t_w, k, role="t2s")
    # agent.py:553,557: in role s2t the learned ability weights are the LARGE model's own kdl_*_weight parameters;
    # [DECISION] a large model built without them (pretraining teacher configs carry no kdl block) uses the small one's
    owner = teacher if hasattr(teacher.bert, "kdl_txt_weight") else None
    Lt = makd_losses(student, t_out, s_out, task, t_rw, s_w, k, role="s2t", weight_owner=owner)
    total_s = k["kd_alpha"] * sum(Ls.values()) + (1 - k["kd_alpha"]) * s_out["loss"].mean()
    total_t = k["t_kd_alpha"] * sum(Lt.values()) + (1 - k["t_kd_alpha"]) * t_out["loss"].mean()  # agent.py:1145
    return total_s, total_t, Ls, Lt, s_out, t_out
