"""Model-config helper: the student / teacher config derivation of pretrain_src/train_r2r_magic.py:125-160
(`teacher_*` / `student_*` keys prefix-stripped onto copies; intermediate = hidden * mlp_ratio,
heads = hidden / 64) for callers that do not go through transformers.PretrainedConfig."""
from types import SimpleNamespace

DEFAULTS = dict(
    pred_head_dropout_prob=0.1, attention_probs_dropout_prob=0.1, hidden_act="gelu", hidden_dropout_prob=0.1,
    hidden_size=768, initializer_range=0.02, intermediate_size=3072, num_l_layers=6, num_x_layers=3,
    num_pano_layers=2, layer_norm_eps=1e-12, max_position_embeddings=514, max_action_steps=100,
    num_attention_heads=12, type_vocab_size=1, update_lang_bert=True, vocab_size=50265, use_lang2visn_attn=True,
    graph_sprels=True, glocal_fuse=True, image_feat_size=768, image_prob_size=1000, angle_feat_size=4,
    obj_feat_size=0, adaptive_pano_fusion=True, cfp_temperature=1.0,
)


def make_config(hidden_size, num_l_layers=6, num_x_layers=3, num_pano_layers=2, mlp_ratio=4, role="student",
                teacher_hidden_size=None, pretrain_tasks=("mlm", "sap"), **over):
    cfg = dict(DEFAULTS)
    cfg.update(hidden_size=hidden_size, num_l_layers=num_l_layers, num_x_layers=num_x_layers,
               num_pano_layers=num_pano_layers, intermediate_size=int(hidden_size * mlp_ratio),
               num_attention_heads=int(hidden_size / 64), role=role, kd=teacher_hidden_size is not None,
               pretrain_tasks=set(pretrain_tasks))
    if teacher_hidden_size is not None:
        cfg["teacher_hidden_size"] = teacher_hidden_size
    cfg.update(over)
    return SimpleNamespace(**cfg)


def split_teacher_student(model_config, knowledge_distillation=True):
    """train_r2r_magic.py:125-160 on a dict/namespace that carries teacher_* and student_* keys."""
    base = dict(vars(model_config)) if not isinstance(model_config, dict) else dict(model_config)

    def derive(prefix, role):
        c = dict(base)
        for k, v in base.items():
            if k.startswith(prefix + "_"):
                c[k[len(prefix) + 1:]] = v
        c["intermediate_size"] = int(c["hidden_size"] * c.get("mlp_ratio", 4))
        c["num_attention_heads"] = int(c["hidden_size"] / 64)
        c["role"], c["kd"] = role, knowledge_distillation
        return c

    teacher = derive("teacher", "teacher") if knowledge_distillation else None
    student = derive("student", "student")
    if teacher is not None:
        for k in list(student.keys()):
            if k.startswith("student_"):
                student["teacher_" + k[8:]] = teacher[k[8:]]
    return (SimpleNamespace(**teacher) if teacher else None), SimpleNamespace(**student)
