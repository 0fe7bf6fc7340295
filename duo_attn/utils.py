from duo_attention_b200.utils import load_attn_pattern, seed_everything, sparsify_attention_heads  # noqa: F401
