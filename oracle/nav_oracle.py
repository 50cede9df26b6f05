"""ORACLE (test infrastructure, NOT product code) -- fp32 PyTorch restatement of the fine-tune / inference path
(SURVEY.md section 8 row f4): `VLNBert` in its three modes and the local -> global logit fusion on viewpoint-id strings.

Only `tests/` may import this file.  The product (`vln-magic_b200/nav.py`) never does.

PARITY UNPINNED (architecture): `map_nav_src/models/model.py` and `models/graph_utils.py` are imported by
/root/reference/map_nav_src/r2r/agent.py:29-30 but are not in the reference tree.  Pinned by reference code and
followed here: the three call signatures and returned tuples (agent.py:797, :885, :964-967), the nav-input dict keys
and the [stop] / [MEM] layout (agent.py:175-245, :289-328 -- `tests/test_nav_host.py` executes that collator source
against ours), the masks (`gmap_masks[:, 1] = False`, `gmap_visited_masks` = [0, 1, 1.., 0..], `vp_nav_masks` =
[1, 0, nav_types == 1]).  The transformer blocks are the pretraining oracle's (magic_oracle.py).
[DECISION]s are the same two as in nav.py's header.
"""
import torch
import torch.nn as nn

from . import magic_oracle as O


class VLNBert(nn.Module):
    def __init__(self, config, role="student"):
        super().__init__()
        if not hasattr(config, "pretrain_tasks"):
            config.pretrain_tasks = ("sap",)
        config.role = role
        self.config = config
        self.vln_bert = O.GlocalTextPathCMTPreTraining(config)

    def forward(self, mode, batch):
        bert = self.vln_bert.bert
        if mode == "language":
            lens = batch["txt_masks"].sum(1)
            x, _, attns = bert.forward_text(batch["txt_ids"], lens)
            return x, attns
        if mode == "panorama":
            type_embed = bert.embeddings.token_type_embeddings.weight[0]
            e, fused, attns = bert.img_embeddings(batch["view_img_fts"], batch["loc_fts"], batch["nav_types"],
                                                  batch["view_lens"], type_embed)
            return e, O.gen_seq_masks(batch["view_lens"], e.shape[1]), fused, attns
        if mode == "navigation":
            return self.forward_navigation(batch)
        raise NotImplementedError(mode)

    def forward_navigation(self, batch):
        m = self.vln_bert
        bert = m.bert
        ge, le = bert.global_encoder, bert.local_encoder
        txt, txt_masks = batch["txt_embeds"], batch["txt_masks"]
        g_in = batch["gmap_img_embeds"] + ge.gmap_step_embeddings(batch["gmap_step_ids"]) + \
            ge.gmap_pos_embeddings(batch["gmap_pos_fts"])
        sprel = None
        if ge.sprel_linear is not None:
            sprel = ge.sprel_linear(batch["gmap_pair_dists"].unsqueeze(3)).squeeze(3).unsqueeze(1)
        g, g_attn = ge.encoder(g_in, txt, batch["gmap_masks"], txt_masks, sprel)
        v_in = batch["vp_img_embeds"] + le.vp_pos_embeddings(batch["vp_pos_fts"])
        v, v_attn = le.encoder(v_in, txt, batch["vp_masks"], txt_masks)
        if m.sap_fuse_linear is None:
            fw = 0.5
        else:
            fw = torch.sigmoid(m.sap_fuse_linear(torch.cat([g[:, 0], v[:, 0]], 1)))
        gl = m.global_sap_head(g).squeeze(2) * fw
        gl = gl.masked_fill(batch["gmap_visited_masks"], float("-inf")).masked_fill(~batch["gmap_masks"], float("-inf"))
        ll = m.local_sap_head(v).squeeze(2) * (1 - fw)
        ll = ll.masked_fill(~batch["vp_nav_masks"], float("-inf"))
        mem = 1 if batch["gmap_vpids"][0][1] is None else 0
        fl = fuse_logits(gl, ll, batch["gmap_vpids"], batch["gmap_visited_masks"], batch["vp_cand_vpids"], 1 + mem)
        slot = 1 if mem else 0
        return {"gmap_embeds": g, "vp_embeds": v, "global_logits": gl, "local_logits": ll, "fused_logits": fl,
                "cls_embeds": g[:, slot] + v[:, slot], "gmap_attns": g_attn, "vp_attns": v_attn}


def fuse_logits(global_logits, local_logits, gmap_vpids, visited_masks, vp_cand_vpids, n_special):
    """DUET-lineage fusion on id strings; slots below `n_special` ([stop], [MEM]) are not candidates."""
    fused = global_logits.clone()
    fused[:, 0] = fused[:, 0] + local_logits[:, 0]
    for b in range(global_logits.shape[0]):
        visited = set(vp for vp, mk in zip(gmap_vpids[b], visited_masks[b].tolist()) if mk)
        tmp, bw = {}, 0
        for j, c in enumerate(vp_cand_vpids[b]):
            if j < n_special:
                continue
            if c in visited:
                bw = bw + local_logits[b, j]
            else:
                tmp[c] = local_logits[b, j]
        for n, vp in enumerate(gmap_vpids[b]):
            if n > 0 and vp not in visited:
                fused[b, n] = fused[b, n] + (tmp[vp] if vp in tmp else bw)
    return fused
