"""The oracle's transformer BLOCKS against the third-party implementations the reference builds its (unshipped)
model from: `transformers` BertLayer / BertSelfAttention / BertSelfOutput / BertIntermediate / BertOutput
(requirements.txt:18; the DUET-lineage model subclasses / copies these, and train_r2r_magic.py:189-208 maps checkpoint
keys onto exactly their parameter names) and `torch.nn.TransformerEncoderLayer` (the panorama encoder, norm_first).
Same weights in, same activations out: the block arithmetic of oracle/magic_oracle.py is pinned; what stays a
[DECISION] is only how the blocks are wired."""
import math

import pytest
import torch

from oracle import magic_oracle as O


def _cfg(h=128):
    return O.make_config(h, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)


def test_bert_layer_matches_transformers():
    tr = pytest.importorskip("transformers")
    from transformers.models.bert.modeling_bert import BertLayer
    c = _cfg()
    hf_cfg = tr.BertConfig(hidden_size=c.hidden_size, num_attention_heads=c.num_attention_heads,
                           intermediate_size=c.intermediate_size, hidden_act="gelu", layer_norm_eps=c.layer_norm_eps,
                           hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    hf_cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    hf = BertLayer(hf_cfg).eval()
    for p in hf.parameters():
        p.data.normal_(0, 0.05)
    mine = O.BertLayer(c).eval()
    missing, unexpected = mine.load_state_dict(hf.state_dict(), strict=False)
    assert not missing and not unexpected, (missing, unexpected)   # identical parameter names and shapes
    B, L = 3, 17
    x = torch.randn(B, L, c.hidden_size)
    lens = torch.tensor([17, 9, 12])
    mask = O.gen_seq_masks(lens, L)
    ext = O.ext_mask(mask)                      # HF "extended attention mask": [B, 1, 1, L] additive
    with torch.no_grad():
        y_hf = hf(x, attention_mask=ext)
        y_hf = y_hf[0] if isinstance(y_hf, (tuple, list)) else y_hf   # (a tuple before transformers 5)
        y, p = mine(x, ext)
    valid = mask[..., None].expand_as(y)
    assert torch.allclose(y[valid], y_hf[valid], rtol=1e-5, atol=1e-6)
    assert torch.allclose(p.sum(-1), torch.ones(B, L), atol=1e-5)   # head-mean of softmaxes


def test_cross_layer_blocks_match_transformers_attention():
    """A METER-style cross layer = HF BertAttention (self) -> HF BertAttention with encoder_hidden_states (cross) -> HF
    BertIntermediate / BertOutput; the oracle's BertCrossLayer with the same weights gives the same output."""
    tr = pytest.importorskip("transformers")
    from transformers.models.bert.modeling_bert import BertAttention, BertIntermediate, BertOutput
    c = _cfg()
    hf_cfg = tr.BertConfig(hidden_size=c.hidden_size, num_attention_heads=c.num_attention_heads,
                           intermediate_size=c.intermediate_size, hidden_act="gelu", layer_norm_eps=c.layer_norm_eps,
                           hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, is_decoder=False)
    hf_cfg._attn_implementation = "eager"
    torch.manual_seed(1)
    try:
        sa, ca = BertAttention(hf_cfg), BertAttention(hf_cfg, is_cross_attention=True)
    except TypeError:
        sa, ca = BertAttention(hf_cfg), BertAttention(hf_cfg)
    inter, outp = BertIntermediate(hf_cfg), BertOutput(hf_cfg)
    for m in (sa, ca, inter, outp):
        m.eval()
        for p in m.parameters():
            p.data.normal_(0, 0.05)
    mine = O.BertCrossLayer(c).eval()
    sd = {}
    for prefix, m in (("attention", sa), ("crossattention", ca), ("intermediate", inter), ("output", outp)):
        for k, v in m.state_dict().items():
            sd[f"{prefix}.{k}"] = v
    missing, unexpected = mine.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    B, Lx, Lc = 2, 11, 19
    x, ctx = torch.randn(B, Lx, c.hidden_size), torch.randn(B, Lc, c.hidden_size)
    xm = O.ext_mask(O.gen_seq_masks(torch.tensor([11, 7]), Lx))
    cm = O.ext_mask(O.gen_seq_masks(torch.tensor([19, 13]), Lc))
    with torch.no_grad():
        a = sa(x, attention_mask=xm)[0]
        try:
            cx = ca(a, attention_mask=cm, encoder_hidden_states=ctx, encoder_attention_mask=cm)[0]
        except TypeError:
            cx = ca(a, encoder_hidden_states=ctx, encoder_attention_mask=cm)[0]
        want = outp(inter(cx), cx)
        got, _ = mine(x, ctx, xm, cm)
    assert torch.allclose(got[:, :7], want[:, :7], rtol=1e-5, atol=1e-6)


def test_pano_layer_matches_torch_transformer_encoder_layer():
    c = _cfg()
    h, H = c.hidden_size, c.num_attention_heads
    torch.manual_seed(2)
    ref = torch.nn.TransformerEncoderLayer(h, H, c.intermediate_size, dropout=0.0, activation="gelu", norm_first=True,
                                           batch_first=True, layer_norm_eps=c.layer_norm_eps).eval()
    for p in ref.parameters():
        p.data.normal_(0, 0.05)
    mine = O.PanoLayer(c).eval()
    missing, unexpected = mine.load_state_dict(ref.state_dict(), strict=False)
    assert not missing and not unexpected, (missing, unexpected)   # nn.TransformerEncoderLayer's own key names
    B, N = 3, 36
    x = torch.randn(B, N, h)
    key_mask = O.gen_seq_masks(torch.tensor([36, 20, 29]), N)
    if hasattr(torch.backends, "mha"):
        torch.backends.mha.set_fastpath_enabled(False)  # the fused inference fast path returns nested / zeroed padding
    with torch.no_grad():
        want = ref(x, src_key_padding_mask=~key_mask)
        got, p = mine(x, key_mask)
    valid = key_mask[..., None].expand_as(got)
    assert torch.allclose(got[valid], want[valid], rtol=1e-5, atol=1e-6)


def test_embeddings_match_transformers_bert_embeddings():
    """Word + position + token-type + LayerNorm of HF BertEmbeddings with absolute position ids arange(L) -- the
    [DECISION] the oracle documents (RoBERTa's padding-offset position ids would shift the table by 2 rows)."""
    tr = pytest.importorskip("transformers")
    from transformers.models.bert.modeling_bert import BertEmbeddings
    c = _cfg()
    hf_cfg = tr.BertConfig(vocab_size=c.vocab_size, hidden_size=c.hidden_size, max_position_embeddings=c.max_position_embeddings,
                           type_vocab_size=c.type_vocab_size, layer_norm_eps=c.layer_norm_eps, hidden_dropout_prob=0.0,
                           pad_token_id=1)
    torch.manual_seed(3)
    hf = BertEmbeddings(hf_cfg).eval()
    mine = O.BertEmbeddings(c).eval()
    sd = {k: v for k, v in hf.state_dict().items() if k in mine.state_dict()}
    missing, _ = mine.load_state_dict(sd, strict=False)
    assert not missing
    ids = torch.randint(3, 50000, (2, 23))
    with torch.no_grad():
        want = hf(input_ids=ids, token_type_ids=torch.zeros_like(ids), position_ids=torch.arange(23)[None])
        got = mine(ids)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)


def test_mlm_head_matches_transformers():
    """`mlm_head.predictions.{transform.dense, transform.LayerNorm, decoder, bias}` (the key names of a DUET-lineage
    pretraining checkpoint) = HF BertOnlyMLMHead: same state dict, same logits."""
    tr = pytest.importorskip("transformers")
    from transformers.models.bert.modeling_bert import BertOnlyMLMHead
    c = _cfg()
    hf_cfg = tr.BertConfig(hidden_size=c.hidden_size, vocab_size=997, hidden_act="gelu", layer_norm_eps=c.layer_norm_eps)
    torch.manual_seed(2)
    hf = BertOnlyMLMHead(hf_cfg).eval()
    for p in hf.parameters():
        p.data.normal_(0, 0.05)
    hf.predictions.decoder.bias = hf.predictions.bias  # the tie a full HF model applies (`_tie_weights`)
    c.vocab_size = 997
    mine = O.BertOnlyMLMHead(c).eval()
    sd = {k: v for k, v in hf.state_dict().items()}
    missing, unexpected = mine.load_state_dict(sd, strict=False)
    # HF ties `decoder.bias` to `predictions.bias` and lists the alias in its state dict; the head keeps the one tensor
    assert not missing and unexpected == ["predictions.decoder.bias"], (missing, unexpected)
    assert torch.equal(sd["predictions.decoder.bias"], sd["predictions.bias"])
    x = torch.randn(11, c.hidden_size)
    with torch.no_grad():
        assert torch.allclose(mine(x), hf(x), rtol=1e-5, atol=1e-6)
